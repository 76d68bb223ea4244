// Fused discrete collision cost/gradient kernel: ONE launch per optimiser step.
//
// Reference semantics: addTimeIntPenaltyParallel (back_end_optimizer.hpp:432-554) with grad_cost_p (hpp:766-824)
// wired into the sample loop exactly as hpp:619-626 wires its swept-volume sibling; PCSmapManager::getPointsInAABB
// (PCSmap_manager.h:148-170); getSDFWithGradWhenRobotAtState (sw_manager.hpp:537-541).
//
// Mapping (B200-first, not the reference's OpenMP-over-samples + critical section):
//   * one warp per pose sample; a CTA owns DISC_WARPS consecutive samples, the hardware CTA scheduler balances
//     the heavy (near-obstacle) stretches of the trajectory across the 148 SMs;
//   * the pose window is read from the BIT-packed occupancy: lane = (x,y) row, one or two 32-bit loads + funnel
//     shift give the row's z-run; a warp prefix sum over popcounts enumerates only the occupied voxels
//     (the reference tests every voxel and heap-allocates a vector per sample, hpp:787);
//   * two warp-level compaction queues keep lanes dense: queue A = voxels inside the body-frame cull box,
//     queue B = voxels whose hinge is active (sdf < safety_hor) and therefore need the 6 extra finite-difference
//     SDF evaluations of getSDFwithGrad1 (Shape.hpp:58-87) — evaluating those only for active voxels is exact
//     because an inactive voxel contributes nothing (hpp:809-821);
//   * per-lane FP64 accumulators, xor-butterfly warp reduction, flatness adjoint, then a deterministic two-level
//     reduction: per-sample partials in HBM, and the LAST warp to finish a piece (ticket counter) sums that piece
//     in ascending sample order — the same order as the serial oracle, independent of scheduling.
#pragma once
#include "isdf_types.cuh"

namespace isdf {

constexpr int DISC_WARPS = 4;
constexpr int DISC_THREADS = DISC_WARPS * 32;
constexpr int QCAP = 64;

struct DiscArgs {
    DevCfg cfg;
    DevGrid grid;
    DevShape shape;
    int N;
    const double *T;       // N
    const double *C;       // 6N x 3 column-major
    double *partial;       // S x PARTIAL_STRIDE
    int *piece_ticket;     // N  (zero on entry, zero on exit)
    int *pieces_done;      // 1  (zero on entry, zero on exit)
    double *piece_cost;    // N
    double *out;           // 19N+1: cost | gradC | gradT
    unsigned long long *pair_counter;  // may be null
    int rank, world;       // this launch evaluates samples s with s % world == rank
};

struct PairAcc { double c, gx, gy, gz, q0, q1, q2, q3; };

// hpp:809-821 for one (pose, voxel) pair
__device__ __forceinline__ void pair_accumulate(const DevCfg &cfg, const rot3 &R, quat4 q, d3 d, double sdf, d3 g, PairAcc &a) {
    double f, df;
    hinge(cfg.safety - sdf, cfg.mu, f, df);
    if (f > 0.0) {
        a.c += f;
        const d3 wg = rot_apply(R, g);
        a.gx += df * wg.x; a.gy += df * wg.y; a.gz += df * wg.z;
        double jq[4];
        quat_pull(q, g, d, jq);
        a.q0 += -df * jq[0]; a.q1 += -df * jq[1]; a.q2 += -df * jq[2]; a.q3 += -df * jq[3];
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ double clampr(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// GridMap3D::getGridIndex after PCSmapManager::projInMap (Gridmap3D.cpp:135-175, PCSmap_manager.h:130-137)
__device__ __forceinline__ int grid_axis_index(double coord, double bmin, double bmax, double res, int size) {
    const double c = clampr(coord, bmin, bmax);
    return clampi((int)floor((c - bmin) / res), 0, size - 1);
}

template <bool MESH>
__global__ void __launch_bounds__(DISC_THREADS) k_discrete(const __grid_constant__ DiscArgs A) {
    __shared__ uint32_t qA[DISC_WARPS][QCAP];
    __shared__ uint32_t qB[DISC_WARPS][QCAP];
    __shared__ double qBs[DISC_WARPS][QCAP];
    __shared__ double stage[DISC_WARPS][PARTIAL_STRIDE];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const DevCfg &cfg = A.cfg;
    const int K = cfg.K, N = A.N;
    const int S = N * (K + 1);
    // local sample index m -> global sample s = rank + world * m
    const int M = (S - A.rank + A.world - 1) / A.world;
    const int m = blockIdx.x * DISC_WARPS + warp;
    if (m >= M) return;
    const int s = A.rank + A.world * m;
    const int i = s / (K + 1), j = s - i * (K + 1);

    // ---- pose of this sample (every lane, redundantly: ~300 FP64 ops, no divergence). Only pos / q / R stay live
    // across the window scan; the epilogue re-derives the rest on one lane to keep register pressure low. -------------
    const double Ti = __ldg(A.T + i);
    const double frac = 1.0 / K;
    const double step = Ti * frac;
    d3 pos; quat4 q; rot3 R;
    {
        double cx[6], cy[6], cz[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            cx[k] = __ldg(A.C + 6 * i + k);
            cy[k] = __ldg(A.C + 6 * N + 6 * i + k);
            cz[k] = __ldg(A.C + 12 * N + 6 * i + k);
        }
        PieceEval pe;
        piece_eval(cx, cy, cz, j * step, pe);
        q = flat_quat_only(cfg.fp, pe.vel, pe.acc);
        pos = pe.pos;
        R = quat_rot(q);
    }

    PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned npairs = 0;

    if (cfg.flags & ISDF_WITH_COLLISION) {
        const DevGrid &G = A.grid;
        const double h = cfg.half_bd;
        const int ix0 = grid_axis_index(pos.x - h, G.bmin[0], G.bmax[0], G.res, G.X);
        const int ix1 = grid_axis_index(pos.x + h, G.bmin[0], G.bmax[0], G.res, G.X);
        const int iy0 = grid_axis_index(pos.y - h, G.bmin[1], G.bmax[1], G.res, G.Y);
        const int iy1 = grid_axis_index(pos.y + h, G.bmin[1], G.bmax[1], G.res, G.Y);
        const int iz0 = grid_axis_index(pos.z - h, G.bmin[2], G.bmax[2], G.res, G.Z);
        const int iz1 = grid_axis_index(pos.z + h, G.bmin[2], G.bmax[2], G.res, G.Z);
        const int ny = iy1 - iy0 + 1;
        const int nrows = (ix1 - ix0 + 1) * ny;
        int nA = 0, nB = 0;  // queue fill (warp-uniform)

        // queue B consumer: finite-difference gradient + accumulation for `cnt` active voxels
        auto drain_B = [&](int cnt) {
            if (lane < cnt) {
                const uint32_t e = qB[warp][lane];
                const double sdf = qBs[warp][lane];
                const d3 ctr = mk3((ix0 + (int)(e & 0xffu) + 0.5) * G.res + G.bmin[0],
                                   (iy0 + (int)((e >> 8) & 0xffu) + 0.5) * G.res + G.bmin[1],
                                   ((int)(e >> 16) + 0.5) * G.res + G.bmin[2]);
                const d3 d = ctr - pos;
                const d3 prel = rot_applyT(R, d);
                const d3 g = shape_grad_analytic(A.shape, prel, sdf);
                pair_accumulate(cfg, R, q, d, sdf, g, acc);
            }
            __syncwarp();
            const int rest = nB - cnt;
            uint32_t e2 = 0; double s2 = 0;
            if (lane < rest) { e2 = qB[warp][cnt + lane]; s2 = qBs[warp][cnt + lane]; }
            __syncwarp();
            if (lane < rest) { qB[warp][lane] = e2; qBs[warp][lane] = s2; }
            __syncwarp();
            nB = rest;
        };
        // queue A consumer: SDF for `cnt` voxels inside the cull box
        auto drain_A = [&](int cnt) {
            bool active = false; uint32_t e = 0; double sdf = 0.0;
            if (lane < cnt) {
                e = qA[warp][lane];
                const d3 ctr = mk3((ix0 + (int)(e & 0xffu) + 0.5) * G.res + G.bmin[0],
                                   (iy0 + (int)((e >> 8) & 0xffu) + 0.5) * G.res + G.bmin[1],
                                   ((int)(e >> 16) + 0.5) * G.res + G.bmin[2]);
                const d3 d = ctr - pos;
                const d3 prel = rot_applyT(R, d);
                if (MESH) {
                    d3 g = mk3(0, 0, 0);
                    sdf = mesh_sdf_grad(A.shape.mesh, prel, cfg.safety, g);
                    pair_accumulate(cfg, R, q, d, sdf, g, acc);
                } else {
                    sdf = shape_sdf_analytic(A.shape, prel);
                    active = (cfg.safety - sdf) > 0.0;
                }
            }
            __syncwarp();
            const int rest = nA - cnt;
            uint32_t e2 = 0;
            if (lane < rest) e2 = qA[warp][cnt + lane];
            __syncwarp();
            if (lane < rest) qA[warp][lane] = e2;
            nA = rest;
            if (!MESH) {
                const unsigned bal = __ballot_sync(0xffffffffu, active);
                if (active) { const int p = nB + __popc(bal & lt_mask); qB[warp][p] = e; qBs[warp][p] = sdf; }
                nB += __popc(bal);
                __syncwarp();
                if (nB >= 32) drain_B(32);
            } else {
                __syncwarp();
            }
        };

        for (int zs = iz0; zs <= iz1; zs += 32) {
            const int nzc = min(32, iz1 - zs + 1);
            const uint32_t zmask = (nzc == 32) ? 0xffffffffu : ((1u << nzc) - 1u);
            const int wz = zs >> 5, sh = zs & 31;
            for (int rb = 0; rb < nrows; rb += 32) {
                const int r = rb + lane;
                uint32_t bits = 0;
                int rx = 0, ry = 0;
                if (r < nrows) {
                    rx = r / ny; ry = r - rx * ny;
                    const uint32_t *row = G.bits + ((size_t)(ix0 + rx) * G.Y + (iy0 + ry)) * G.Zw;
                    const uint32_t lo = __ldg(row + wz);
                    const uint32_t hi = (sh != 0 && wz + 1 < G.Zw) ? __ldg(row + wz + 1) : 0u;
                    bits = __funnelshift_r(lo, hi, sh) & zmask;
                }
                const int cnt = __popc(bits);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                for (int base = 0; base < total; base += 32) {
                    const int kk = base + lane;
                    // source lane = number of lanes whose inclusive count is <= kk (binary search over the scan)
                    int src = 0;
#pragma unroll
                    for (int stp = 16; stp > 0; stp >>= 1) {
                        const int probe = src + stp - 1;
                        const int v = __shfl_sync(0xffffffffu, incl, probe & 31);
                        if (probe < 32 && v <= kk) src += stp;
                    }
                    src = min(src, 31);
                    const int s_incl = __shfl_sync(0xffffffffu, incl, src);
                    const int s_cnt = __shfl_sync(0xffffffffu, cnt, src);
                    const uint32_t s_bits = __shfl_sync(0xffffffffu, bits, src);
                    const int s_rx = __shfl_sync(0xffffffffu, rx, src), s_ry = __shfl_sync(0xffffffffu, ry, src);
                    bool pass = false; uint32_t entry = 0;
                    if (kk < total) {
                        const int nth = kk - (s_incl - s_cnt);                 // 0-based rank inside the source row
                        const int bz = __fns(s_bits, 0, nth + 1);              // position of the nth set bit
                        const int lz = (zs - 0) + bz;                          // absolute z index
                        const d3 ctr = mk3((ix0 + s_rx + 0.5) * G.res + G.bmin[0], (iy0 + s_ry + 0.5) * G.res + G.bmin[1],
                                           (lz + 0.5) * G.res + G.bmin[2]);
                        const d3 prel = rot_applyT(R, ctr - pos);
                        pass = !(fabs(prel.x) > h || fabs(prel.y) > h || fabs(prel.z) > h);  // hpp:800
                        if (pass) npairs++;
                        if (MESH && pass) {
                            // outside the mesh AABB inflated by safety_hor => sdf >= safety_hor => hinge inactive (exact skip)
                            const DevMesh &Mh = A.shape.mesh;
                            const double sf = cfg.safety;
                            if (prel.x < Mh.blo[0] - sf || prel.x > Mh.bhi[0] + sf || prel.y < Mh.blo[1] - sf || prel.y > Mh.bhi[1] + sf ||
                                prel.z < Mh.blo[2] - sf || prel.z > Mh.bhi[2] + sf) pass = false;
                        }
                        entry = (uint32_t)s_rx | ((uint32_t)s_ry << 8) | ((uint32_t)lz << 16);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) qA[warp][nA + __popc(bal & lt_mask)] = entry;
                    nA += __popc(bal);
                    __syncwarp();
                    if (nA >= 32) drain_A(32);
                }
            }
        }
        if (nA > 0) drain_A(nA);
        if (!MESH && nB > 0) drain_B(nB);
    }

    // ---- reduce the pair sums over the warp ----------------------------------------------------------------------
    const double costp = warp_sum(acc.c);
    const d3 gp = mk3(warp_sum(acc.gx), warp_sum(acc.gy), warp_sum(acc.gz));
    const double gq0 = warp_sum(acc.q0), gq1 = warp_sum(acc.q1), gq2 = warp_sum(acc.q2), gq3 = warp_sum(acc.q3);
    if (A.pair_counter) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, npairs);
        if (lane == 0 && tot) atomicAdd(A.pair_counter, (unsigned long long)tot);
    }

    // ---- sample epilogue (hpp:505-551) ----------------------------------------------------------------------------
    if (lane == 0) {
        double cx[6], cy[6], cz[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            cx[k] = __ldg(A.C + 6 * i + k);
            cy[k] = __ldg(A.C + 6 * N + 6 * i + k);
            cz[k] = __ldg(A.C + 12 * N + 6 * i + k);
        }
        PieceEval pe;
        piece_eval(cx, cy, cz, j * step, pe);
        FlatState fs;
        flat_state(cfg.fp, pe.vel, pe.acc, pe.jer, fs);
        const d3 omg = flat_omega(fs);
        double pena = 0.0;
        d3 gradVel = mk3(0, 0, 0), gradOmg = mk3(0, 0, 0), gradPos = mk3(0, 0, 0);
        double gradQuat[4] = {0, 0, 0, 0};
        if (cfg.flags & ISDF_WITH_DYNAMICS) {
            const double cos_theta = 1.0 - 2.0 * (q.x * q.x + q.y * q.y);
            double f, df;
            if (hinge(dot3(pe.vel, pe.vel) - cfg.vmax2, cfg.mu, f, df)) {
                const double sc = cfg.wv * df * 2.0;
                gradVel = mk3(sc * pe.vel.x, sc * pe.vel.y, sc * pe.vel.z); pena += cfg.wv * f;
            }
            if (hinge(dot3(omg, omg) - cfg.omgmax2, cfg.mu, f, df)) {
                const double sc = cfg.womg * df * 2.0;
                gradOmg = mk3(sc * omg.x, sc * omg.y, sc * omg.z); pena += cfg.womg * f;
            }
            if (hinge(acos(cos_theta) - cfg.thetamax, cfg.mu, f, df)) {
                const double sc = cfg.wtheta * df / sqrt(1.0 - cos_theta * cos_theta) * 4.0;
                gradQuat[1] += sc * q.x; gradQuat[2] += sc * q.y; pena += cfg.wtheta * f;
            }
        }
        if (costp > 0.0) {  // grad_cost_p returns (costp > 0), hpp:823
            gradPos = mk3(cfg.wp * gp.x, cfg.wp * gp.y, cfg.wp * gp.z);
            gradQuat[0] += cfg.wp * gq0; gradQuat[1] += cfg.wp * gq1; gradQuat[2] += cfg.wp * gq2; gradQuat[3] += cfg.wp * gq3;
            pena += cfg.wp * costp;
        }
        d3 gV, gA, gJ;
        flat_adjoint(cfg.fp, fs, pe.vel, pe.acc, gradQuat, gradOmg, gradVel, gV, gA, gJ);
        const d3 gP = gradPos;
        const double node = (j == 0 || j == K) ? 0.5 : 1.0;
        const double alpha = j * frac;
        const double w = node * step;
        double *st = stage[warp];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            st[k] = (pe.b0[k] * gP.x + pe.b1[k] * gV.x + pe.b2[k] * gA.x + pe.b3[k] * gJ.x) * w;
            st[6 + k] = (pe.b0[k] * gP.y + pe.b1[k] * gV.y + pe.b2[k] * gA.y + pe.b3[k] * gJ.y) * w;
            st[12 + k] = (pe.b0[k] * gP.z + pe.b1[k] * gV.z + pe.b2[k] * gA.z + pe.b3[k] * gJ.z) * w;
        }
        st[18] = (dot3(gP, pe.vel) + dot3(gV, pe.acc) + dot3(gA, pe.jer) + dot3(gJ, pe.sna)) * alpha * node * step + node * frac * pena;
        st[19] = node * step * pena;
    }
    __syncwarp();
    if (lane < PARTIAL_STRIDE) A.partial[(size_t)s * PARTIAL_STRIDE + lane] = stage[warp][lane];

    // ---- deterministic reduction: last warp of a piece sums it in ascending-sample order ------------------------
    __threadfence();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(A.piece_ticket + i, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    // number of this rank's samples in piece i
    const int first_s = i * (K + 1), last_s = first_s + K;
    int f0 = first_s + ((A.rank - first_s) % A.world + A.world) % A.world;  // first local sample >= first_s
    const int local_cnt = (f0 > last_s) ? 0 : ((last_s - f0) / A.world + 1);
    if (ticket != local_cnt - 1) return;
    __threadfence();
    if (lane < PARTIAL_STRIDE) {
        double sum = 0.0;
        for (int ss = f0; ss <= last_s; ss += A.world) sum += __ldcg(A.partial + (size_t)ss * PARTIAL_STRIDE + lane);
        if (lane < 18) { const int ax = lane / 6, k = lane - 6 * ax; A.out[1 + (size_t)ax * 6 * N + 6 * i + k] = sum; }
        else if (lane == 18) A.out[1 + 18 * N + i] = sum;
        else A.piece_cost[i] = sum;
    }
    if (lane == 0) A.piece_ticket[i] = 0;
    __threadfence();
    int done = 0;
    if (lane == 0) done = atomicAdd(A.pieces_done, 1);
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done != N - 1) return;
    __threadfence();
    if (lane == 0) {
        double c = 0.0;
        for (int p = 0; p < N; p++) c += __ldcg(A.piece_cost + p);
        A.out[0] = c;
        *A.pieces_done = 0;
    }
}

}  // namespace isdf
