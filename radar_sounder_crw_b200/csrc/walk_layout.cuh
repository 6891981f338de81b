// walk_layout.cuh -- workspace layouts shared by the generic (walk_f32.cu) and small-N (walk_small.cu) walk kernels.
#pragma once
#include "common.cuh"

namespace crw {

// layout of the `saved` workspace (floats)
struct WalkLayout {
    size_t invn, A, S, Sp, L, R, G, part, total;
    int B, T, N, C;
    __host__ __device__ WalkLayout(int B_, int T_, int N_, int C_) : B(B_), T(T_), N(N_), C(C_) {
        const size_t nn = (size_t)N * N, bt1 = (size_t)B * (T - 1);
        size_t o = 0;
        invn = o; o += align_up((size_t)B * T * N, 64);
        A = o;    o += align_up(bt1 * nn, 64);
        S = o;    o += align_up(bt1 * nn, 64);
        Sp = o;   o += align_up(bt1 * nn, 64);
        L = o;    o += align_up(bt1 * nn, 64);   // L_k, k = 0..T-2
        R = o;    o += align_up(bt1 * nn, 64);   // R_k, k = 0..T-2 (k = 0 unused)
        G = o;    o += align_up(bt1 * nn, 64);   // G_k = rowsoftmax(M_k) - I, k = 1..T-2
        part = o; o += align_up(bt1, 64);        // loss partials [B][T-1]
        total = o;
    }
    __host__ __device__ size_t mat(size_t base, int b, int t) const { return base + ((size_t)b * (T - 1) + t) * N * N; }
};

// ------------------------------------------------------------------------------------------
// backward scratch layout (floats): dL, dR [B][T-1][N][N] (index k), dS, dSp [B][T-1][N][N] (index t)
// ------------------------------------------------------------------------------------------
struct BwdLayout {
    size_t dL, dR, dS, dSp, dAw, total;
    __host__ __device__ BwdLayout(int B, int T, int N) {
        const size_t m = align_up((size_t)B * (T - 1) * N * N, 64);
        dL = 0; dR = m; dS = 2 * m; dSp = 3 * m; dAw = 4 * m; total = 5 * m;
    }
};


}  // namespace crw
