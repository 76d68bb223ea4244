// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C wrapper around the ONE reference file that compiles in this environment:
//   /root/reference/src/utils/include/igl/FastWindingNumberForSoups.h  (UT_SolidAngle<float,float>, :6197-7250),
// the arithmetic behind igl::fast_winding_number(fwn_bvh, 2.0, p) used at Shape.cpp:110,121,131,144.
// It is included from where it lies (never copied); the built oracle/_ref/libref_fwn.so is the known-answer
// source for winding numbers (kind "reference"). The libigl wrapper around it lives only in lib/libigl.a
// (libigl 2.4.0); per its published source it casts V and queries to float and calls init(..., order) —
// reproduced below.
#include <igl/FastWindingNumberForSoups.h>
#include <vector>
#include <cmath>

namespace igl { unsigned int default_num_threads(unsigned int) { return 1; } }  // body absent from the header tree

using namespace igl::FastWindingNumber::HDK_Sample;

extern "C" {
struct RefFwn {
    std::vector<UT_Vector3T<float>> P;
    std::vector<int> tris;
    UT_SolidAngle<float, float> sa;
};
void *ref_fwn_create(const double *V, int nV, const int *F, int nF, int order) {
    RefFwn *r = new RefFwn();
    r->P.resize(nV);
    for (int i = 0; i < nV; i++) { r->P[i][0] = (float)V[3 * i]; r->P[i][1] = (float)V[3 * i + 1]; r->P[i][2] = (float)V[3 * i + 2]; }
    r->tris.assign(F, F + 3 * (size_t)nF);
    r->sa.init(nF, r->tris.data(), nV, r->P.data(), order);
    return r;
}
void ref_fwn_query(void *h, const double *q, int n, double accuracy_scale, double *w_out) {
    RefFwn *r = (RefFwn *)h;
    for (int i = 0; i < n; i++) {
        UT_Vector3T<float> p; p[0] = (float)q[3 * i]; p[1] = (float)q[3 * i + 1]; p[2] = (float)q[3 * i + 2];
        w_out[i] = (double)(r->sa.computeSolidAngle(p, (float)accuracy_scale) / (4.0 * M_PI));
    }
}
void ref_fwn_destroy(void *h) { delete (RefFwn *)h; }
}
