// sm_100a kernels for the per-edge half of the LM inner loop:
//   errors_kernel        SparseOptimizer::computeActiveErrors + activeRobustChi2   (sparse_optimizer.cpp:63-116)
//   build_pl_kernel      linearizeOplus + constructQuadraticForm for (pose, landmark) edges: Hpl block, Hll/b_l segment sums
//   pose_accum/reduce    the pose half of the same quadratic form, as a deterministic pose-sorted second pass
//   build_pp_kernel      the same for (pose, pose) edges                               (base_binary_edge.hpp:62-137)
//   update_kernel        SparseOptimizer::update -> oplusImpl                          (sparse_optimizer.cpp:441-455)
// Layout: edges of one type form one SoA "edge set", pose-landmark sets sorted by (landmark, pose) so that the Hpl
// blocks of one landmark are contiguous (== the reference's _HplCCS order) and Hll/b_l reduce inside a CTA.
#include "edge_math.cuh"
#include "kernels.hpp"

namespace g2ocu {

template <int ET> struct Role { static constexpr bool PL = false; static constexpr int PS = 0; };
template <> struct Role<G2OCU_EDGE_SE2_POINT_XY> { static constexpr bool PL = true; static constexpr int PS = 0; };
template <> struct Role<G2OCU_EDGE_PROJECT_XYZ2UV> { static constexpr bool PL = true; static constexpr int PS = 1; };
template <> struct Role<G2OCU_EDGE_SE3_PROJECT_XYZ> { static constexpr bool PL = true; static constexpr int PS = 1; };
template <> struct Role<G2OCU_EDGE_BAL> { static constexpr bool PL = true; static constexpr int PS = 0; };

constexpr int kThreads = 128;

template <int E> G2D void loadInfo(const EdgeSetDev& s, int i, double* Om) {
  if (s.infoMode == 0) {
#pragma unroll
    for (int k = 0; k < E * E; ++k) Om[k] = ((k % (E + 1)) == 0) ? 1.0 : 0.0;
  } else {
    const double* p = s.info + (s.infoMode == 2 ? (size_t)i * E * E : 0);
#pragma unroll
    for (int k = 0; k < E * E; ++k) Om[k] = __ldg(p + k);
  }
}

// loads estimates / measurement / parameters of edge i (kernel order) into registers
template <int ET> G2D void loadEdge(const EdgeSetDev& s, const SystemDev& sys, int i, int slot0, int slot1, double* x0, double* x1, double* z, double* prm) {
  using T = EdgeT<ET>;
  const double* b0 = (Role<ET>::PL && Role<ET>::PS != 0) ? sys.lmEst : sys.poseEst;
  const double* b1 = (Role<ET>::PL && Role<ET>::PS != 1) ? sys.lmEst : sys.poseEst;
  const double* p0 = b0 + (size_t)slot0 * T::S0; const double* p1 = b1 + (size_t)slot1 * T::S1;
#pragma unroll
  for (int k = 0; k < T::S0; ++k) x0[k] = __ldg(p0 + k);
#pragma unroll
  for (int k = 0; k < T::S1; ++k) x1[k] = __ldg(p1 + k);
  const double* pz = s.meas + (size_t)i * T::M;
#pragma unroll
  for (int k = 0; k < T::M; ++k) z[k] = __ldg(pz + k);
  if (T::NP > 0) {
    const double* pp = s.prm + (s.prmMode == 2 ? (size_t)i * T::NP : 0);
#pragma unroll
    for (int k = 0; k < (T::NP > 0 ? T::NP : 1); ++k) prm[k] = __ldg(pp + k);
  }
}

// chi2 = e^T Ω e and the robust weight; returns rho0, sets w = rho1
template <int E> G2D double robustWeight(const EdgeSetDev& s, int i, const double* e, const double* Om, double& chi2, double& w) {
  chi2 = 0;
#pragma unroll
  for (int j = 0; j < E; ++j) { double t = 0;
#pragma unroll
    for (int k = 0; k < E; ++k) t += Om[j + E * k] * e[k];
    chi2 += e[j] * t; }
  double rho0 = chi2; w = 1.0;
  if (s.kernelMode) {
    const int kind = s.kernelMode == 2 ? s.kernelKind[i] : s.kKind;
    const double delta = s.kernelMode == 2 ? s.kernelDelta[i] : s.kDelta;
    if (kind) robustify_dev(kind, delta, chi2, rho0, w);
  }
  return rho0;
}

G2D double warpSum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum (fixed shuffle tree + fixed warp order); result valid on thread 0
template <int NT> G2D double blockSum(double v, double* sm /* NT/32 doubles */) {
  v = warpSum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) r += sm[k];
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
template <int ET> __global__ void __launch_bounds__(kThreads) errors_kernel(EdgeSetDev s, SystemDev sys, double* partial, double* errOut, const int64_t* errOff) {
  using T = EdgeT<ET>;
  __shared__ double sm[kThreads / 32];
  double rho0 = 0, chi2 = 0;
  // grid-stride over the edges: a fixed edge -> thread assignment and fixed reduction trees keep chi2 bit-reproducible from run to run
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < s.n; i += gridDim.x * kThreads) {
    double x0[T::S0], x1[T::S1], z[T::M], prm[T::NP > 0 ? T::NP : 1], e[T::E], Om[T::E * T::E], w, c2;
    loadEdge<ET>(s, sys, i, s.slot0[i], s.slot1[i], x0, x1, z, prm);
    T::template eval<false>(x0, x1, z, prm, e, nullptr, nullptr);
    loadInfo<T::E>(s, i, Om);
    rho0 += robustWeight<T::E>(s, i, e, Om, c2, w);
    chi2 += c2;
    if (errOut) { double* o = errOut + errOff[s.pos[i]];
#pragma unroll
      for (int k = 0; k < T::E; ++k) o[k] = e[k]; }
  }
  const double a = blockSum<kThreads>(rho0, sm);
  const double b = blockSum<kThreads>(chi2, sm);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b; }
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* partial, int nBlocks, int stride, double* out) {
  __shared__ double sm[8];
  for (int c = 0; c < stride; ++c) {
    double v = 0;
    for (int k = threadIdx.x; k < nBlocks; k += 256) v += partial[(size_t)k * stride + c];
    const double r = blockSum<256>(v, sm);
    if (threadIdx.x == 0) out[c] += r;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// (pose, landmark) edges: thread per edge, edges sorted by landmark.
template <int ET> __global__ void __launch_bounds__(kThreads) build_pl_kernel(EdgeSetDev s, SystemDev sys) {
  using T = EdgeT<ET>;
  constexpr int PS = Role<ET>::PS;
  constexpr int P = PS == 0 ? T::D0 : T::D1, L = PS == 0 ? T::D1 : T::D0, E = T::E;
  constexpr int PLn = P * L, NL = L * (L + 1) / 2 + L;
  __shared__ double sB[kThreads * PLn];
  __shared__ double sC[kThreads * NL];
  __shared__ int sLm[kThreads];
  __shared__ int sBlk[kThreads];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kThreads + tid;
  int lm = -1, blk = -1;
  if (i < s.n) {
    const int s0 = s.slot0[i], s1 = s.slot1[i];
    const int lslot = PS == 0 ? s1 : s0;
    double x0[T::S0], x1[T::S1], z[T::M], prm[T::NP > 0 ? T::NP : 1], e[E], J0[E * T::D0], J1[E * T::D1], Om[E * E], w, chi2;
    loadEdge<ET>(s, sys, i, s0, s1, x0, x1, z, prm);
    T::template eval<true>(x0, x1, z, prm, e, J0, J1);
    loadInfo<E>(s, i, Om);
    robustWeight<E>(s, i, e, Om, chi2, w);
    const double* Jp = PS == 0 ? J0 : J1; const double* Jl = PS == 0 ? J1 : J0;
    double OJl[E * L], wr[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < E; ++k) t += Om[r + E * k] * e[k];
      wr[r] = -w * t;
#pragma unroll
      for (int c = 0; c < L; ++c) { double u = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) u += Om[r + E * k] * Jl[k + E * c];
        OJl[r + E * c] = w * u; }
    }
    if (lslot < sys.numLandmarks) {
      lm = lslot;
      int q = 0;
#pragma unroll
      for (int c = 0; c < L; ++c)
#pragma unroll
        for (int r = 0; r <= c; ++r) { double u = 0;
#pragma unroll
          for (int k = 0; k < E; ++k) u += Jl[k + E * r] * OJl[k + E * c];
          sC[tid * NL + q++] = u; }
#pragma unroll
      for (int r = 0; r < L; ++r) { double u = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) u += Jl[k + E * r] * wr[k];
        sC[tid * NL + q++] = u; }
    }
    blk = s.block[i];
    if (blk >= 0) {
#pragma unroll
      for (int c = 0; c < L; ++c)
#pragma unroll
        for (int r = 0; r < P; ++r) { double u = 0;
#pragma unroll
          for (int k = 0; k < E; ++k) u += Jp[k + E * r] * OJl[k + E * c];
          sB[tid * PLn + r + P * c] = u; }
    }
  }
  sLm[tid] = lm; sBlk[tid] = blk;
  __syncthreads();
  // coalesced write of the Hpl blocks of this CTA (contiguous when the blocks are)
  for (int t = tid; t < kThreads * PLn; t += kThreads) {
    const int j = t / PLn, k = t - j * PLn, b = sBlk[j];
    if (b >= 0) {
      if (sys.hplShared) atomicAdd(sys.Hpl + (size_t)b * PLn + k, sB[t]);
      else sys.Hpl[(size_t)b * PLn + k] = sB[t];
    }
  }
  // landmark segments: the first thread of each run sums the run (<= 128 long) and adds it to Hll / b_l
  if (lm >= 0 && (tid == 0 || sLm[tid - 1] != lm)) {
    double acc[NL];
#pragma unroll
    for (int q = 0; q < NL; ++q) acc[q] = sC[tid * NL + q];
    for (int t = tid + 1; t < kThreads && sLm[t] == lm; ++t) {
#pragma unroll
      for (int q = 0; q < NL; ++q) acc[q] += sC[t * NL + q];
    }
    double* H = sys.Hll + (size_t)lm * L * L;
    int q = 0;
#pragma unroll
    for (int c = 0; c < L; ++c)
#pragma unroll
      for (int r = 0; r <= c; ++r) { atomicAdd(H + r + L * c, acc[q]); if (r != c) atomicAdd(H + c + L * r, acc[q]); ++q; }
    double* bl = sys.b + (size_t)sys.numPoses * P + (size_t)lm * L;
#pragma unroll
    for (int r = 0; r < L; ++r) atomicAdd(bl + r, acc[q++]);
  }
}

// pose half: one CTA per chunk (<= kChunk edges, all of one pose), register accumulation, deterministic tree reduce
template <int ET> __global__ void __launch_bounds__(kThreads) pose_accum_kernel(EdgeSetDev s, SystemDev sys) {
  using T = EdgeT<ET>;
  constexpr int PS = Role<ET>::PS;
  constexpr int P = PS == 0 ? T::D0 : T::D1, E = T::E;
  constexpr int SP = PS == 0 ? T::S0 : T::S1, SL = PS == 0 ? T::S1 : T::S0;
  constexpr int NU = P * (P + 1) / 2, NV = NU + P;
  __shared__ double sm[(kThreads / 32) * NV];
  const int chunk = blockIdx.x;
  const int pose = s.chunkPose[chunk], begin = s.chunkBegin[chunk], end = s.chunkEnd[chunk];
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = 0;
  double xp[SP];
#pragma unroll
  for (int k = 0; k < SP; ++k) xp[k] = __ldg(sys.poseEst + (size_t)pose * SP + k);
  for (int j = begin + threadIdx.x; j < end; j += kThreads) {
    const int i = s.byPose[j];
    const int lslot = PS == 0 ? s.slot1[i] : s.slot0[i];
    double xl[SL], z[T::M], prm[T::NP > 0 ? T::NP : 1], e[E], J0[E * T::D0], J1[E * T::D1], Om[E * E], w, chi2;
#pragma unroll
    for (int k = 0; k < SL; ++k) xl[k] = __ldg(sys.lmEst + (size_t)lslot * SL + k);
#pragma unroll
    for (int k = 0; k < T::M; ++k) z[k] = __ldg(s.meas + (size_t)i * T::M + k);
    if (T::NP > 0) {
      const double* pp = s.prm + (s.prmMode == 2 ? (size_t)i * T::NP : 0);
#pragma unroll
      for (int k = 0; k < (T::NP > 0 ? T::NP : 1); ++k) prm[k] = __ldg(pp + k);
    }
    if (PS == 0) T::template eval<true>(xp, xl, z, prm, e, J0, J1); else T::template eval<true>(xl, xp, z, prm, e, J0, J1);
    loadInfo<E>(s, i, Om);
    robustWeight<E>(s, i, e, Om, chi2, w);
    const double* Jp = PS == 0 ? J0 : J1;
    double OJp[E * P], wr[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < E; ++k) t += Om[r + E * k] * e[k];
      wr[r] = -w * t;
#pragma unroll
      for (int c = 0; c < P; ++c) { double u = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) u += Om[r + E * k] * Jp[k + E * c];
        OJp[r + E * c] = w * u; }
    }
    int q = 0;
#pragma unroll
    for (int c = 0; c < P; ++c)
#pragma unroll
      for (int r = 0; r <= c; ++r) { double u = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) u += Jp[k + E * r] * OJp[k + E * c];
        acc[q++] += u; }
#pragma unroll
    for (int r = 0; r < P; ++r) { double u = 0;
#pragma unroll
      for (int k = 0; k < E; ++k) u += Jp[k + E * r] * wr[k];
      acc[q++] += u; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) { const double v = warpSum(acc[q]); if (lane == 0) sm[wid * NV + q] = v; }
  __syncthreads();
  for (int q = threadIdx.x; q < NV; q += kThreads) {
    double v = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) v += sm[k * NV + q];
    s.partial[(size_t)chunk * NV + q] = v;
  }
}

template <int P> __global__ void pose_reduce_kernel(EdgeSetDev s, SystemDev sys) {
  constexpr int NU = P * (P + 1) / 2, NV = NU + P;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= sys.numPoses * NV) return;
  const int pose = t / NV, q = t - pose * NV;
  const int c0 = s.poseChunkPtr[pose], c1 = s.poseChunkPtr[pose + 1];
  if (c0 == c1) return;
  double v = 0;
  for (int c = c0; c < c1; ++c) v += s.partial[(size_t)c * NV + q];
  if (q < NU) {
    int col = 0; while ((col + 1) * (col + 2) / 2 <= q) ++col;
    const int row = q - col * (col + 1) / 2;
    double* H = sys.Hpp + (size_t)sys.hppDiag[pose] * P * P;
    H[row + P * col] += v;
    if (row != col) H[col + P * row] += v;
  } else sys.b[(size_t)pose * P + (q - NU)] += v;
}

// (pose, pose) edges: thread per edge, atomics on the diagonal blocks (degree is small), plain store off-diagonal
template <int ET> __global__ void __launch_bounds__(kThreads) build_pp_kernel(EdgeSetDev s, SystemDev sys) {
  using T = EdgeT<ET>;
  constexpr int E = T::E, P = T::D0;
  static_assert(T::D0 == T::D1, "pose-pose edges connect equal-sized blocks");
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= s.n) return;
  const int s0 = s.slot0[i], s1 = s.slot1[i];
  double x0[T::S0], x1[T::S1], z[T::M], prm[1], e[E], J0[E * P], J1[E * P], Om[E * E], w, chi2;
  loadEdge<ET>(s, sys, i, s0, s1, x0, x1, z, prm);
  T::template eval<true>(x0, x1, z, prm, e, J0, J1);
  loadInfo<E>(s, i, Om);
  robustWeight<E>(s, i, e, Om, chi2, w);
  double wr[E];
#pragma unroll
  for (int r = 0; r < E; ++r) { double t = 0;
#pragma unroll
    for (int k = 0; k < E; ++k) t += Om[r + E * k] * e[k];
    wr[r] = -w * t; }
#pragma unroll
  for (int k = 0; k < E * E; ++k) Om[k] *= w;
  const bool f0 = s0 < sys.numPoses, f1 = s1 < sys.numPoses;
  double OJ[E * P];
  auto omegaTimes = [&](const double* J) {
#pragma unroll
    for (int c = 0; c < P; ++c)
#pragma unroll
      for (int r = 0; r < E; ++r) { double u = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) u += Om[r + E * k] * J[k + E * c];
        OJ[r + E * c] = u; }
  };
  auto diagAdd = [&](const double* J, int slot) {
    double* H = sys.Hpp + (size_t)sys.hppDiag[slot] * P * P;
#pragma unroll
    for (int c = 0; c < P; ++c)
#pragma unroll
      for (int r = 0; r <= c; ++r) { double u = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) u += J[k + E * r] * OJ[k + E * c];
        atomicAdd(H + r + P * c, u); if (r != c) atomicAdd(H + c + P * r, u); }
    double* bb = sys.b + (size_t)slot * P;
#pragma unroll
    for (int r = 0; r < P; ++r) { double u = 0;
#pragma unroll
      for (int k = 0; k < E; ++k) u += J[k + E * r] * wr[k];
      atomicAdd(bb + r, u); }
  };
  const int blk = s.block[i];
  if (f0) {
    omegaTimes(J0);
    diagAdd(J0, s0);
    if (blk >= 0 && s.transposed[i]) {        // block(row = idx1, col = idx0) = J1^T Ω J0
      double* H = sys.Hpp + (size_t)blk * P * P;
#pragma unroll
      for (int c = 0; c < P; ++c)
#pragma unroll
        for (int r = 0; r < P; ++r) { double u = 0;
#pragma unroll
          for (int k = 0; k < E; ++k) u += J1[k + E * r] * OJ[k + E * c];
          if (sys.hppShared) atomicAdd(H + r + P * c, u); else H[r + P * c] = u; }
    }
  }
  if (f1) {
    omegaTimes(J1);
    diagAdd(J1, s1);
    if (blk >= 0 && !s.transposed[i]) {       // block(row = idx0, col = idx1) = J0^T Ω J1
      double* H = sys.Hpp + (size_t)blk * P * P;
#pragma unroll
      for (int c = 0; c < P; ++c)
#pragma unroll
        for (int r = 0; r < P; ++r) { double u = 0;
#pragma unroll
          for (int k = 0; k < E; ++k) u += J0[k + E * r] * OJ[k + E * c];
          if (sys.hppShared) atomicAdd(H + r + P * c, u); else H[r + P * c] = u; }
    }
  }
}

template <int ET> __global__ void __launch_bounds__(kThreads) jacobian_dump_kernel(EdgeSetDev s, SystemDev sys, double* out, const int64_t* off) {
  using T = EdgeT<ET>;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= s.n) return;
  double x0[T::S0], x1[T::S1], z[T::M], prm[T::NP > 0 ? T::NP : 1], e[T::E], J0[T::E * T::D0], J1[T::E * T::D1];
  loadEdge<ET>(s, sys, i, s.slot0[i], s.slot1[i], x0, x1, z, prm);
  T::template eval<true>(x0, x1, z, prm, e, J0, J1);
  double* o = out + off[s.pos[i]];
#pragma unroll
  for (int k = 0; k < T::E * T::D0; ++k) o[k] = J0[k];
#pragma unroll
  for (int k = 0; k < T::E * T::D1; ++k) o[T::E * T::D0 + k] = J1[k];
}

template <int VT> __global__ void update_kernel(double* est, int* counters, const double* x, int nFree) {
  using V = VertexT<VT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFree) return;
  double v[V::S], u[V::D];
#pragma unroll
  for (int k = 0; k < V::S; ++k) v[k] = est[(size_t)i * V::S + k];
#pragma unroll
  for (int k = 0; k < V::D; ++k) u[k] = x[(size_t)i * V::D + k];
  int cnt = counters ? counters[i] : 0;
  V::oplus(v, u, &cnt);
  if (counters) counters[i] = cnt;
#pragma unroll
  for (int k = 0; k < V::S; ++k) est[(size_t)i * V::S + k] = v[k];
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
#define FOR_EDGE_TYPE(et, CALL)                                                      \
  switch (et) {                                                                      \
    case G2OCU_EDGE_SE2: { CALL(G2OCU_EDGE_SE2) break; }                             \
    case G2OCU_EDGE_SE2_POINT_XY: { CALL(G2OCU_EDGE_SE2_POINT_XY) break; }           \
    case G2OCU_EDGE_SE3: { CALL(G2OCU_EDGE_SE3) break; }                             \
    case G2OCU_EDGE_SE3_EXPMAP: { CALL(G2OCU_EDGE_SE3_EXPMAP) break; }               \
    case G2OCU_EDGE_PROJECT_XYZ2UV: { CALL(G2OCU_EDGE_PROJECT_XYZ2UV) break; }       \
    case G2OCU_EDGE_SE3_PROJECT_XYZ: { CALL(G2OCU_EDGE_SE3_PROJECT_XYZ) break; }     \
    case G2OCU_EDGE_BAL: { CALL(G2OCU_EDGE_BAL) break; }                             \
    default: break;                                                                  \
  }

static int errorGrid(int n) { const int nb = (n + kThreads - 1) / kThreads; return nb < 148 * 16 ? nb : 148 * 16; }   // at most 16 CTAs per SM, grid-stride beyond
int errorScratchDoubles(int n) { return 2 * errorGrid(n) + 2; }

void launchErrors(const EdgeSetDev& s, const SystemDev& sys, double* scratch, double* out2, double* errOut, const int64_t* errOff, cudaStream_t st, int64_t* launches) {
  if (s.n == 0) return;
  const int nb = errorGrid(s.n);
#define CALL(ETV) errors_kernel<ETV><<<nb, kThreads, 0, st>>>(s, sys, scratch, errOut, errOff);
  FOR_EDGE_TYPE(s.etype, CALL)
#undef CALL
  sum_partials_kernel<<<1, 256, 0, st>>>(scratch, nb, 2, out2);
  *launches += 2;
}

template <int ET> static void buildPL(const EdgeSetDev& s, const SystemDev& sys, cudaStream_t st, int64_t* launches) {
  using T = EdgeT<ET>;
  constexpr int P = Role<ET>::PS == 0 ? T::D0 : T::D1;
  constexpr int NV = P * (P + 1) / 2 + P;
  const int nb = (s.n + kThreads - 1) / kThreads;
  build_pl_kernel<ET><<<nb, kThreads, 0, st>>>(s, sys);
  *launches += 1;
  if (s.nChunks > 0) {
    pose_accum_kernel<ET><<<s.nChunks, kThreads, 0, st>>>(s, sys);
    const int tot = sys.numPoses * NV;
    pose_reduce_kernel<P><<<(tot + 255) / 256, 256, 0, st>>>(s, sys);
    *launches += 2;
  }
}
template <int ET> static void buildPP(const EdgeSetDev& s, const SystemDev& sys, cudaStream_t st, int64_t* launches) {
  const int nb = (s.n + kThreads - 1) / kThreads;
  build_pp_kernel<ET><<<nb, kThreads, 0, st>>>(s, sys);
  *launches += 1;
}

void launchBuild(const EdgeSetDev& s, const SystemDev& sys, cudaStream_t st, int64_t* launches) {
  if (s.n == 0) return;
  switch (s.etype) {
    case G2OCU_EDGE_SE2: buildPP<G2OCU_EDGE_SE2>(s, sys, st, launches); break;
    case G2OCU_EDGE_SE3: buildPP<G2OCU_EDGE_SE3>(s, sys, st, launches); break;
    case G2OCU_EDGE_SE3_EXPMAP: buildPP<G2OCU_EDGE_SE3_EXPMAP>(s, sys, st, launches); break;
    case G2OCU_EDGE_SE2_POINT_XY: buildPL<G2OCU_EDGE_SE2_POINT_XY>(s, sys, st, launches); break;
    case G2OCU_EDGE_PROJECT_XYZ2UV: buildPL<G2OCU_EDGE_PROJECT_XYZ2UV>(s, sys, st, launches); break;
    case G2OCU_EDGE_SE3_PROJECT_XYZ: buildPL<G2OCU_EDGE_SE3_PROJECT_XYZ>(s, sys, st, launches); break;
    case G2OCU_EDGE_BAL: buildPL<G2OCU_EDGE_BAL>(s, sys, st, launches); break;
    default: break;
  }
}

void launchJacobianDump(const EdgeSetDev& s, const SystemDev& sys, double* jacOut, const int64_t* jacOff, cudaStream_t st, int64_t* launches) {
  if (s.n == 0) return;
  const int nb = (s.n + kThreads - 1) / kThreads;
#define CALL(ETV) jacobian_dump_kernel<ETV><<<nb, kThreads, 0, st>>>(s, sys, jacOut, jacOff);
  FOR_EDGE_TYPE(s.etype, CALL)
#undef CALL
  *launches += 1;
}

void launchUpdate(int vtype, double* est, double*, int* counters, const double* x, int nFree, cudaStream_t st, int64_t* launches) {
  if (nFree == 0) return;
  const int nb = (nFree + 255) / 256;
  switch (vtype) {
    case G2OCU_VERTEX_SE2: update_kernel<G2OCU_VERTEX_SE2><<<nb, 256, 0, st>>>(est, nullptr, x, nFree); break;
    case G2OCU_VERTEX_POINT_XY: update_kernel<G2OCU_VERTEX_POINT_XY><<<nb, 256, 0, st>>>(est, nullptr, x, nFree); break;
    case G2OCU_VERTEX_SE3: update_kernel<G2OCU_VERTEX_SE3><<<nb, 256, 0, st>>>(est, counters, x, nFree); break;
    case G2OCU_VERTEX_SE3_EXPMAP: update_kernel<G2OCU_VERTEX_SE3_EXPMAP><<<nb, 256, 0, st>>>(est, nullptr, x, nFree); break;
    case G2OCU_VERTEX_POINT_XYZ: update_kernel<G2OCU_VERTEX_POINT_XYZ><<<nb, 256, 0, st>>>(est, nullptr, x, nFree); break;
    case G2OCU_VERTEX_CAM_BAL: update_kernel<G2OCU_VERTEX_CAM_BAL><<<nb, 256, 0, st>>>(est, nullptr, x, nFree); break;
    case G2OCU_VERTEX_POINT_BAL: update_kernel<G2OCU_VERTEX_POINT_BAL><<<nb, 256, 0, st>>>(est, nullptr, x, nFree); break;
    default: return;
  }
  *launches += 1;
}

}  // namespace g2ocu
