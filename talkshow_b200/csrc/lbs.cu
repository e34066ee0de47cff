// talkshow_b200 — batched SMPL-X evaluation (linear blend skinning) for all frames of a batch at once.
//
// Reference: scripts/demo.py:122-152 (get_vertices) and data_utils/get_j.py:20-51 (get_joints) call the third-party
// `smplx` body model ONE FRAME AT A TIME in float64 on the CPU.  This is the same algorithm (smplx 0.1.28:
// SMPLX.forward + lbs.lbs, restated in oracle/smplx_oracle.py) for F frames per call in fp32:
//   1. per frame: full pose (+ hand mean), 55 Rodrigues rotations, blend-shape coefficients
//      c = [expression(100) | (R_1..R_54 - I)(486)], joints J = J0 + JS.[betas | expression], kinematic chain ->
//      relative transforms A[55][3x4] and the posed joints;
//   2. ONE GEMM  v_posed[F, 3V] = c[F, 586] . dirs[586, 3V] + v_base   (v_base = v_template + shapedirs.betas);
//   3. skinning: per (frame, vertex)  T = sum_j w[v][j] A[j],  vertex = T . [v_posed, 1];
//   4. extra joints (selected vertices) and the 51 static face landmarks (barycentric on the landmark triangles).
#include "pixelcnn.h"

#include <cmath>
#include <memory>

namespace ts {

constexpr int LBS_J = 55;
constexpr int LBS_NB = 300, LBS_NE = 100, LBS_NP = 486;
constexpr int LBS_KC = 588;          // expression + pose-feature coefficients, padded to a multiple of 4
constexpr int LBS_MAXF = 4096;       // frames per GEMM chunk (v_posed workspace: 4096 x 3V floats)

struct SmplxModel {
  int V = 0, nextra = 0, nlmk = 0;
  Layer dirs;            // W [3V][588]: columns 0..99 expression dirs, 100..585 pose dirs; bias unused
  Layer beta_dirs;       // W [3V][300]
  float* v_template = nullptr;   // [3V]
  float* J0 = nullptr;           // [165]  J_regressor . v_template
  float* JS = nullptr;           // [165][400]  J_regressor . shapedirs
  float* wT = nullptr;           // [55][V] skinning weights, joint-major
  float* pose_mean = nullptr;    // [165]
  int* parents = nullptr;        // [55]
  int* extra_idx = nullptr;      // [nextra]
  int* lmk_tri = nullptr;        // [nlmk][3] vertex ids of the landmark triangles
  float* lmk_bary = nullptr;     // [nlmk][3]
};

void smplx_destroy(ts_engine* e) {
  delete (SmplxModel*)e->smplx;
  e->smplx = nullptr;
}

static int* upload_i(ts_engine* e, const std::vector<int>& h) {
  int* d = (int*)e->dmalloc(h.size() * sizeof(int));
  if (d) TS_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
  return d;
}

// pose265 layout of the reference's call sites (demo.py:129-138): jaw 0:3 | leye 3:6 | reye 6:9 | global 9:12 |
// body 12:75 | left hand 75:120 | right hand 120:165 | expression 165:265  ->  SMPL-X joint order
// [global, body x21, jaw, leye, reye, left hand x15, right hand x15]
__device__ __forceinline__ int lbs_src(int joint) {
  if (joint == 0) return 9;
  if (joint <= 21) return 12 + (joint - 1) * 3;
  if (joint <= 24) return (joint - 22) * 3;
  return 75 + (joint - 25) * 3;
}

__global__ void __launch_bounds__(64) lbs_pose_kernel(const float* __restrict__ pose, const float* __restrict__ betas,
                                                      const float* __restrict__ pose_mean, const float* __restrict__ J0,
                                                      const float* __restrict__ JS, const int* __restrict__ parents, float* coef,
                                                      float* A, float* joints, int F, int use_expr, int njoints_out) {
  __shared__ float R[LBS_J][9];
  __shared__ float Jl[LBS_J * 3];
  __shared__ float G[LBS_J][12];
  __shared__ float shp[LBS_NB + LBS_NE];
  const int f = blockIdx.x, t = threadIdx.x;
  const float* p = pose + (size_t)f * 265;
  for (int i = t; i < LBS_NB + LBS_NE; i += 64)
    shp[i] = i < LBS_NB ? (betas ? betas[i] : 0.f) : (use_expr ? p[165 + i - LBS_NB] : 0.f);
  float* c = coef + (size_t)f * LBS_KC;
  for (int i = t; i < LBS_NE; i += 64) c[i] = use_expr ? p[165 + i] : 0.f;
  if (t < LBS_KC - LBS_NE - LBS_NP) c[LBS_NE + LBS_NP + t] = 0.f;
  if (t < LBS_J) {
    // batch_rodrigues (smplx/lbs.py): angle = |v + 1e-8| (epsilon on the vector), axis = v / angle
    const int s = lbs_src(t);
    const float x = p[s] + pose_mean[3 * t], y = p[s + 1] + pose_mean[3 * t + 1], z = p[s + 2] + pose_mean[3 * t + 2];
    const float ex = x + 1e-8f, ey = y + 1e-8f, ez = z + 1e-8f;
    const float ang = sqrtf(ex * ex + ey * ey + ez * ez);
    const float rx = x / ang, ry = y / ang, rz = z / ang;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    const float oc = 1.f - cs;
    // R = I + sin K + (1 - cos) K^2,  K = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]]
    R[t][0] = 1.f + oc * (-rz * rz - ry * ry);
    R[t][1] = -sn * rz + oc * (rx * ry);
    R[t][2] = sn * ry + oc * (rx * rz);
    R[t][3] = sn * rz + oc * (rx * ry);
    R[t][4] = 1.f + oc * (-rz * rz - rx * rx);
    R[t][5] = -sn * rx + oc * (ry * rz);
    R[t][6] = -sn * ry + oc * (rx * rz);
    R[t][7] = sn * rx + oc * (ry * rz);
    R[t][8] = 1.f + oc * (-ry * ry - rx * rx);
    if (t >= 1)
      for (int k = 0; k < 9; ++k) c[LBS_NE + (t - 1) * 9 + k] = R[t][k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
  }
  __syncthreads();
  for (int i = t; i < LBS_J * 3; i += 64) {
    const float* js = JS + (size_t)i * (LBS_NB + LBS_NE);
    float a = J0[i];
    for (int k = 0; k < LBS_NB + LBS_NE; ++k) a = fmaf(js[k], shp[k], a);
    Jl[i] = a;
  }
  __syncthreads();
  if (t == 0) {
    // batch_rigid_transform: G_i = G_parent . [R_i | J_i - J_parent]
    for (int i = 0; i < LBS_J; ++i) {
      const int pa = parents[i];
      float tr[3], L[12];
      for (int k = 0; k < 3; ++k) tr[k] = Jl[3 * i + k] - (i ? Jl[3 * pa + k] : 0.f);
      for (int r = 0; r < 3; ++r) {
        L[4 * r] = R[i][3 * r]; L[4 * r + 1] = R[i][3 * r + 1]; L[4 * r + 2] = R[i][3 * r + 2]; L[4 * r + 3] = tr[r];
      }
      if (i == 0) {
        for (int k = 0; k < 12; ++k) G[0][k] = L[k];
      } else {
        for (int r = 0; r < 3; ++r)
          for (int cidx = 0; cidx < 4; ++cidx) {
            float a = G[pa][4 * r] * L[cidx] + G[pa][4 * r + 1] * L[4 + cidx] + G[pa][4 * r + 2] * L[8 + cidx];
            if (cidx == 3) a += G[pa][4 * r + 3];
            G[i][4 * r + cidx] = a;
          }
      }
    }
  }
  __syncthreads();
  if (t < LBS_J) {
    // posed joint = translation of G; relative transform A = G with t - G[:3,:3] J (removes the rest-pose joint)
    float* a = A + ((size_t)f * LBS_J + t) * 12;
    float* jo = joints + ((size_t)f * njoints_out + t) * 3;
    for (int r = 0; r < 3; ++r) {
      const float g0 = G[t][4 * r], g1 = G[t][4 * r + 1], g2 = G[t][4 * r + 2], g3 = G[t][4 * r + 3];
      a[4 * r] = g0; a[4 * r + 1] = g1; a[4 * r + 2] = g2;
      a[4 * r + 3] = g3 - (g0 * Jl[3 * t] + g1 * Jl[3 * t + 1] + g2 * Jl[3 * t + 2]);
      jo[r] = g3;
    }
  }
}

__global__ void __launch_bounds__(256) lbs_skin_kernel(const float* __restrict__ vposed, const float* __restrict__ A,
                                                       const float* __restrict__ wT, float* __restrict__ verts, int V, int F) {
  __shared__ float As[LBS_J * 12];
  const int f = blockIdx.y;
  for (int i = threadIdx.x; i < LBS_J * 12; i += 256) As[i] = A[(size_t)f * LBS_J * 12 + i];
  __syncthreads();
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= V) return;
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = 0.f;
  for (int j = 0; j < LBS_J; ++j) {
    const float w = wT[(size_t)j * V + v];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = fmaf(w, As[j * 12 + k], T[k]);
  }
  const float* vp = vposed + ((size_t)f * V + v) * 3;
  const float x = vp[0], y = vp[1], z = vp[2];
  float* o = verts + ((size_t)f * V + v) * 3;
  o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
  o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
  o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
}

__global__ void lbs_extra_kernel(const float* __restrict__ verts, const int* __restrict__ extra_idx, const int* __restrict__ lmk_tri,
                                 const float* __restrict__ lmk_bary, float* joints, int V, int nextra, int nlmk, int njoints_out) {
  const int f = blockIdx.x, t = threadIdx.x;
  const float* vf = verts + (size_t)f * V * 3;
  float* jo = joints + ((size_t)f * njoints_out + LBS_J) * 3;
  if (t < nextra) {
    const int v = extra_idx[t];
    for (int k = 0; k < 3; ++k) jo[3 * t + k] = vf[3 * v + k];
  } else if (t < nextra + nlmk) {
    const int l = t - nextra;
    for (int k = 0; k < 3; ++k) {
      float a = 0.f;
      for (int q = 0; q < 3; ++q) a += vf[3 * lmk_tri[3 * l + q] + k] * lmk_bary[3 * l + q];
      jo[3 * t + k] = a;
    }
  }
}

}  // namespace ts

using namespace ts;

// tensors (host, see oracle/smplx_oracle.py for the shapes): v_template [V,3], shapedirs [V,3,400], posedirs [486,3V],
// J_regressor [55,V], lbs_weights [V,55], pose_mean [165] f32; parents [55], faces [Fc,3], lmk_faces_idx [L],
// extra_joint_idx [E] int64; lmk_bary_coords [L,3] f32.
extern "C" int ts_load_smplx(ts_engine* e, const ts_tensor* tensors, int n) {
  TS_API_BEGIN(e)
  Ckpt ck(tensors, n);
  const ts_tensor* vt = ck.get("v_template");
  if (vt->ndim != 2 || vt->shape[1] != 3) fail(TS_ERR_MISSING, "smplx: v_template must be [V,3]");
  const int V = (int)vt->shape[0];
  const int N3 = 3 * V;
  const float* vtemp = ck.f32("v_template", {V, 3});
  const float* sd = ck.f32("shapedirs", {V, 3, LBS_NB + LBS_NE});
  const float* pd = ck.f32("posedirs", {LBS_NP, N3});
  const float* jr = ck.f32("J_regressor", {LBS_J, V});
  const float* lw = ck.f32("lbs_weights", {V, LBS_J});
  const float* pm = ck.f32("pose_mean", {LBS_J * 3});
  auto i64 = [&](const char* k, int* len) {
    const ts_tensor* t = ck.get(k);
    if (t->dtype != 1) fail(TS_ERR_MISSING, "smplx: tensor '%s' must be int64", k);
    int64_t nn = 1;
    for (int i = 0; i < t->ndim; ++i) nn *= t->shape[i];
    *len = (int)nn;
    return (const int64_t*)t->data;
  };
  int np = 0, nf = 0, nl = 0, ne = 0;
  const int64_t* par = i64("parents", &np);
  const int64_t* faces = i64("faces", &nf);
  const int64_t* lmi = i64("lmk_faces_idx", &nl);
  const int64_t* exi = i64("extra_joint_idx", &ne);
  if (np != LBS_J) fail(TS_ERR_MISSING, "smplx: %d parents (55 expected)", np);
  const float* bary = ck.f32("lmk_bary_coords", {nl, 3});
  if (ne + nl > 256) fail(TS_ERR_UNSUPPORTED, "smplx: %d extra joints + landmarks (> 256)", ne + nl);
  LoadScope scope(e, "smplx");
  std::unique_ptr<SmplxModel> M(new SmplxModel());
  M->V = V; M->nextra = ne; M->nlmk = nl;
  {
    std::vector<float> W((size_t)N3 * LBS_KC, 0.f);     // [3V][588]
    for (int i = 0; i < N3; ++i) {
      float* row = W.data() + (size_t)i * LBS_KC;
      for (int k = 0; k < LBS_NE; ++k) row[k] = sd[(size_t)i * (LBS_NB + LBS_NE) + LBS_NB + k];
      for (int k = 0; k < LBS_NP; ++k) row[LBS_NE + k] = pd[(size_t)k * N3 + i];
    }
    M->dirs.N = N3; M->dirs.K = LBS_KC; M->dirs.W = e->upload(W);
    std::vector<float> Wb((size_t)N3 * LBS_NB);
    for (int i = 0; i < N3; ++i)
      for (int k = 0; k < LBS_NB; ++k) Wb[(size_t)i * LBS_NB + k] = sd[(size_t)i * (LBS_NB + LBS_NE) + k];
    M->beta_dirs.N = N3; M->beta_dirs.K = LBS_NB; M->beta_dirs.W = e->upload(Wb);
  }
  M->v_template = e->upload(std::vector<float>(vtemp, vtemp + N3));
  {
    // J0 = J_regressor . v_template, JS = J_regressor . shapedirs (float64 accumulation, rounded once)
    std::vector<float> J0(LBS_J * 3), JS((size_t)LBS_J * 3 * (LBS_NB + LBS_NE));
    std::vector<double> acc((size_t)3 * (LBS_NB + LBS_NE));
    for (int j = 0; j < LBS_J; ++j) {
      std::fill(acc.begin(), acc.end(), 0.0);
      double j0[3] = {0, 0, 0};
      for (int v = 0; v < V; ++v) {
        const double w = jr[(size_t)j * V + v];
        if (w == 0.0) continue;
        for (int k = 0; k < 3; ++k) {
          j0[k] += w * vtemp[3 * v + k];
          const float* s = sd + ((size_t)v * 3 + k) * (LBS_NB + LBS_NE);
          double* a = acc.data() + (size_t)k * (LBS_NB + LBS_NE);
          for (int c = 0; c < LBS_NB + LBS_NE; ++c) a[c] += w * s[c];
        }
      }
      for (int k = 0; k < 3; ++k) {
        J0[3 * j + k] = (float)j0[k];
        for (int c = 0; c < LBS_NB + LBS_NE; ++c) JS[((size_t)3 * j + k) * (LBS_NB + LBS_NE) + c] = (float)acc[(size_t)k * (LBS_NB + LBS_NE) + c];
      }
    }
    M->J0 = e->upload(J0);
    M->JS = e->upload(JS);
  }
  {
    std::vector<float> wT((size_t)LBS_J * V);
    for (int v = 0; v < V; ++v)
      for (int j = 0; j < LBS_J; ++j) wT[(size_t)j * V + v] = lw[(size_t)v * LBS_J + j];
    M->wT = e->upload(wT);
  }
  M->pose_mean = e->upload(std::vector<float>(pm, pm + LBS_J * 3));
  std::vector<int> pi(LBS_J), ei(ne), tri((size_t)nl * 3);
  for (int i = 0; i < LBS_J; ++i) {
    pi[i] = (int)par[i];
    if (i && (pi[i] < 0 || pi[i] >= i)) fail(TS_ERR_MISSING, "smplx: parents[%d] = %d is not an earlier joint", i, pi[i]);
  }
  for (int i = 0; i < ne; ++i) {
    ei[i] = (int)exi[i];
    if (ei[i] < 0 || ei[i] >= V) fail(TS_ERR_MISSING, "smplx: extra joint vertex %d out of range", ei[i]);
  }
  for (int l = 0; l < nl; ++l) {
    const int64_t fidx = lmi[l];
    if (fidx < 0 || fidx * 3 + 2 >= nf) fail(TS_ERR_MISSING, "smplx: landmark face %lld out of range", (long long)fidx);
    for (int q = 0; q < 3; ++q) {
      tri[3 * l + q] = (int)faces[fidx * 3 + q];
      if (tri[3 * l + q] < 0 || tri[3 * l + q] >= V) fail(TS_ERR_MISSING, "smplx: landmark vertex out of range");
    }
  }
  M->parents = upload_i(e, pi);
  M->extra_idx = upload_i(e, ei);
  M->lmk_tri = upload_i(e, tri);
  M->lmk_bary = e->upload(std::vector<float>(bary, bary + (size_t)nl * 3));
  smplx_destroy(e);
  e->smplx = M.release();
  scope.commit();
  TS_API_END(e)
}

extern "C" int ts_smplx_dims(ts_engine* e, int* V, int* njoints) {
  TS_API_BEGIN(e)
  if (!e->smplx) fail(TS_ERR_NOT_LOADED, "smplx model not loaded");
  const SmplxModel* M = (const SmplxModel*)e->smplx;
  if (V) *V = M->V;
  if (njoints) *njoints = LBS_J + M->nextra + M->nlmk;
  TS_API_END(e)
}

// poses [F,265] (device), betas [300] (device) or NULL (zeros) -> vertices [F,V,3] (may be NULL: joints only still needs
// the vertices internally), joints [F, 55 + extra + landmarks, 3] (may be NULL).
extern "C" int ts_smplx_forward(ts_engine* e, const float* poses, const float* betas, int use_expression, float* vertices,
                                float* joints, int F, void* stream) {
  TS_API_BEGIN(e)
  if (!e->smplx) fail(TS_ERR_NOT_LOADED, "smplx model not loaded");
  if (e->host_only) fail(TS_ERR_UNSUPPORTED, "host-only engine cannot execute");
  if (F <= 0 || !poses) fail(TS_ERR_INVALID, "ts_smplx_forward: F=%d", F);
  const SmplxModel* M = (const SmplxModel*)e->smplx;
  cudaStream_t s = (cudaStream_t)stream;
  const int V = M->V, N3 = 3 * V, NJ = LBS_J + M->nextra + M->nlmk;
  auto body = [&] {
    const int fc = std::min(F, LBS_MAXF);
    float* vbase = e->ws.alloc<float>(N3);
    float* coef = e->ws.alloc<float>((size_t)fc * LBS_KC);
    float* A = e->ws.alloc<float>((size_t)fc * LBS_J * 12);
    float* vposed = e->ws.alloc<float>((size_t)fc * N3);
    float* jtmp = joints ? nullptr : e->ws.alloc<float>((size_t)fc * NJ * 3);
    float* vtmp = vertices ? nullptr : e->ws.alloc<float>((size_t)fc * N3);
    if (e->ws.sizing) return;
    // v_base = v_template + shapedirs[:, :300] . betas  (one row GEMM; betas == NULL -> v_template)
    if (betas) {
      GemmP g;
      g.A = betas; g.W = M->beta_dirs.W; g.bias = M->v_template; g.C = vbase;
      g.M = 1; g.N = N3; g.K = LBS_NB; g.mper = 1; g.a_bs = LBS_NB; g.a_rs = LBS_NB; g.kc = LBS_NB; g.a_ts = LBS_NB;
      g.c_bs = N3; g.c_rs = N3; g.ldw = LBS_NB;
      launch_gemm(e, g, s);
    } else {
      TS_CUDA(cudaMemcpyAsync(vbase, M->v_template, (size_t)N3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    for (int f0 = 0; f0 < F; f0 += fc) {
      const int nf = std::min(fc, F - f0);
      float* jout = joints ? joints + (size_t)f0 * NJ * 3 : jtmp;
      float* vout = vertices ? vertices + (size_t)f0 * N3 : vtmp;
      lbs_pose_kernel<<<nf, 64, 0, s>>>(poses + (size_t)f0 * 265, betas, M->pose_mean, M->J0, M->JS, M->parents, coef, A, jout, nf,
                                        use_expression, NJ);
      e->launches++;
      GemmP g;
      g.A = coef; g.W = M->dirs.W; g.bias = vbase; g.C = vposed;
      g.M = nf; g.N = N3; g.K = LBS_KC; g.mper = nf; g.a_bs = 0; g.a_rs = LBS_KC; g.kc = LBS_KC; g.a_ts = LBS_KC;
      g.c_bs = 0; g.c_rs = N3; g.ldw = LBS_KC;
      launch_gemm(e, g, s);
      lbs_skin_kernel<<<dim3(cdiv(V, 256), nf), 256, 0, s>>>(vposed, A, M->wT, vout, V, nf);
      e->launches++;
      lbs_extra_kernel<<<nf, 256, 0, s>>>(vout, M->extra_idx, M->lmk_tri, M->lmk_bary, jout, V, M->nextra, M->nlmk, NJ);
      e->launches++;
      TS_CUDA(cudaGetLastError());
    }
  };
  e->ws.begin_sizing();
  body();
  const size_t need = e->ws.need;
  e->ws.buf.ensure(need + 256);
  e->ws.begin(need);
  body();
  TS_API_END(e)
}
