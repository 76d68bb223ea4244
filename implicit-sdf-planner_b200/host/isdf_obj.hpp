// Wavefront OBJ -> (V, F) for isdf_set_shape_mesh: the product's counterpart of igl::read_triangle_mesh as the reference's Generalshape
// constructor uses it (src/utils/src/Shape.cpp:36: `igl::read_triangle_mesh(objpath, V, F)`, then the poly_params pre-transform :38-50,
// which isdf_set_shape_mesh applies itself). libigl 2.4.0 is shipped to the reference as a prebuilt archive only; per its published
// readOBJ / read_triangle_mesh: `v x y z [w]` vertices, `f` records with 1-based (or negative = relative) indices in the forms
// i, i/t, i/t/n, i//n; vt / vn / everything else ignored for (V, F); faces with more than three corners are split into a triangle fan.
// Every OBJ the reference ships (src/plan_manager/shapes/*.obj) is plain `v` + triangular `f`.
#pragma once
#include "isdf.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace isdf_host {

inline bool read_obj(const char *path, std::vector<double> &V, std::vector<int32_t> &F, std::string &err) {
    V.clear(); F.clear();
    FILE *fp = std::fopen(path, "r");
    if (!fp) { err = std::string("cannot open ") + path; return false; }
    std::vector<char> line(1 << 16);
    std::vector<long> corners;
    long lineno = 0;
    bool ok = true;
    while (ok && std::fgets(line.data(), (int)line.size(), fp)) {
        lineno++;
        char *s = line.data();
        while (*s == ' ' || *s == '\t') s++;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            char *e = s + 1;
            double xyz[3];
            for (int k = 0; k < 3 && ok; k++) {
                char *n = nullptr;
                xyz[k] = std::strtod(e, &n);
                if (n == e) { err = "bad vertex on line " + std::to_string(lineno); ok = false; }
                e = n;
            }
            if (ok) { V.push_back(xyz[0]); V.push_back(xyz[1]); V.push_back(xyz[2]); }
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            corners.clear();
            char *e = s + 1;
            for (;;) {
                while (*e == ' ' || *e == '\t') e++;
                if (*e == '\0' || *e == '\n' || *e == '\r' || *e == '#') break;
                char *n = nullptr;
                long idx = std::strtol(e, &n, 10);
                if (n == e) { err = "bad face on line " + std::to_string(lineno); ok = false; break; }
                const long nv = (long)(V.size() / 3);
                if (idx < 0) idx = nv + idx + 1;                       // relative index
                if (idx < 1 || idx > nv) { err = "face index out of range on line " + std::to_string(lineno); ok = false; break; }
                corners.push_back(idx - 1);
                e = n;
                while (*e != '\0' && *e != ' ' && *e != '\t' && *e != '\n' && *e != '\r') e++;   // skip /t/n
            }
            if (ok && corners.size() < 3) { err = "face with fewer than three corners on line " + std::to_string(lineno); ok = false; }
            for (size_t k = 1; ok && k + 1 < corners.size(); k++) {    // triangle fan
                F.push_back((int32_t)corners[0]); F.push_back((int32_t)corners[k]); F.push_back((int32_t)corners[k + 1]);
            }
        }
    }
    std::fclose(fp);
    if (ok && (V.empty() || F.empty())) { err = "no vertices or faces in " + std::string(path); ok = false; }
    return ok;
}

// Generalshape(objpath, poly_params) (Shape.cpp:27-50) in one call: read the OBJ, hand it to the device with its pre-transform.
inline int set_shape_obj(isdf_ctx *ctx, const char *path, const double *poly_params6, std::string *err_out = nullptr) {
    std::vector<double> V; std::vector<int32_t> F; std::string err;
    if (!read_obj(path, V, F, err)) { if (err_out) *err_out = err; return ISDF_ERR_INVALID; }
    const int r = isdf_set_shape_mesh(ctx, V.data(), (int)(V.size() / 3), F.data(), (int)(F.size() / 3), poly_params6);
    if (r != ISDF_OK && err_out) *err_out = isdf_last_error();
    return r;
}

}  // namespace isdf_host
