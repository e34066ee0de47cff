// talkshow_b200 — audio front-end on the device: the transform chain of the
// reference's get_mfcc_ta (data_utils/utils.py:148-177) = torchaudio Resample(sr0 -> 22 kHz,
// sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) -> MFCC(n_mfcc 64, n_fft 2048, hop 734,
// hann, center/reflect, power 2, 256 HTK mel bands, AmplitudeToDB(top_db 80), DCT-II ortho).
// The STFT is a GEMM against a [2050 x 2048] DFT matrix (cos | -sin), the mel projection and the DCT are
// GEMMs too (fp32 FFMA kernel); tables are built in double precision at first use.
#include <cmath>

#include "convstack.h"

namespace ts {

constexpr int MF_NFFT = 2048, MF_NFREQ = 1025, MF_NFREQ_PAD = 1028, MF_NMEL = 256, MF_NMFCC = 64, MF_SR = 22000;
constexpr int MF_HOP = 734;

struct MfccTables {
  Layer dft, mel, dct;            // GEMM weights [2050][2048], [256][1028], [64][256]
  float* window = nullptr;        // hann(2048), periodic
  // resample kernel cache for one (orig, new) pair
  int rs_orig = 0, rs_new = 0, rs_width = 0, rs_kw = 0;
  float* rs_kernel = nullptr;     // [new][2*width+orig]
};


static Layer make_layer(ts_engine* e, const std::vector<float>& W, int N, int K) {
  Layer L;
  L.N = N; L.K = K; L.taps = 1; L.cin = K;
  L.W = e->upload(W);
  return L;
}

static MfccTables* tables(ts_engine* e) {
  if (e->mfcc_tables) return (MfccTables*)e->mfcc_tables;   // owned by the engine (device memory dies with it)
  MfccTables* T = new MfccTables();
  e->mfcc_tables = T;
  const double PI = 3.14159265358979323846;
  {  // DFT matrix rows: k < 1025 -> cos(2 pi k n / N); 1025 + k -> -sin(2 pi k n / N)
    std::vector<float> W((size_t)2 * MF_NFREQ * MF_NFFT);
    for (int k = 0; k < MF_NFREQ; ++k)
      for (int n = 0; n < MF_NFFT; ++n) {
        const double a = 2.0 * PI * (double)(((long long)k * n) % MF_NFFT) / MF_NFFT;
        W[(size_t)k * MF_NFFT + n] = (float)std::cos(a);
        W[(size_t)(MF_NFREQ + k) * MF_NFFT + n] = (float)(-std::sin(a));
      }
    T->dft = make_layer(e, W, 2 * MF_NFREQ, MF_NFFT);
  }
  {  // HTK mel filterbank, torchaudio.functional.melscale_fbanks(1025, 0, sr/2, 256, sr, norm=None, 'htk')
    auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
    auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
    const double fmax = MF_SR / 2, mmin = hz2mel(0.0), mmax = hz2mel(fmax);
    std::vector<double> fpts(MF_NMEL + 2);
    for (int i = 0; i < MF_NMEL + 2; ++i) fpts[i] = mel2hz(mmin + (mmax - mmin) * i / (MF_NMEL + 1));
    std::vector<float> W((size_t)MF_NMEL * MF_NFREQ_PAD, 0.f);
    for (int k = 0; k < MF_NFREQ; ++k) {
      const double f = fmax * k / (MF_NFREQ - 1);
      for (int m = 0; m < MF_NMEL; ++m) {
        const double down = (f - fpts[m]) / (fpts[m + 1] - fpts[m]);
        const double up = (fpts[m + 2] - f) / (fpts[m + 2] - fpts[m + 1]);
        const double v = std::max(0.0, std::min(down, up));
        W[(size_t)m * MF_NFREQ_PAD + k] = (float)v;
      }
    }
    T->mel = make_layer(e, W, MF_NMEL, MF_NFREQ_PAD);
  }
  {  // DCT-II, norm='ortho' (torchaudio.functional.create_dct(64, 256, 'ortho'))
    std::vector<float> W((size_t)MF_NMFCC * MF_NMEL);
    for (int k = 0; k < MF_NMFCC; ++k)
      for (int n = 0; n < MF_NMEL; ++n) {
        double v = std::cos(PI / MF_NMEL * (n + 0.5) * k) * std::sqrt(2.0 / MF_NMEL);
        if (k == 0) v *= 1.0 / std::sqrt(2.0);
        W[(size_t)k * MF_NMEL + n] = (float)v;
      }
    T->dct = make_layer(e, W, MF_NMFCC, MF_NMEL);
  }
  {
    std::vector<float> w(MF_NFFT);
    for (int n = 0; n < MF_NFFT; ++n) w[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / MF_NFFT));
    T->window = e->upload(w);
  }
  return T;
}

void mfcc_destroy(ts_engine* e) {
  delete (MfccTables*)e->mfcc_tables;
  e->mfcc_tables = nullptr;
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// torchaudio.functional._get_sinc_resample_kernel (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99)
static void resample_kernel_for(ts_engine* e, MfccTables* T, int sr0) {
  const int g = gcd_i(sr0, MF_SR), orig = sr0 / g, nw = MF_SR / g;
  if (T->rs_orig == orig && T->rs_new == nw) return;
  const double PI = 3.14159265358979323846, lpw = 6.0, rolloff = 0.99;
  const double base = std::min(orig, nw) * rolloff;
  const int width = (int)std::ceil(lpw * orig / base);
  const int kw = 2 * width + orig;
  std::vector<float> K((size_t)nw * kw);
  for (int p = 0; p < nw; ++p)
    for (int j = 0; j < kw; ++j) {
      double t = (-(double)p / nw + (double)(j - width) / orig) * base;
      t = std::max(-lpw, std::min(lpw, t));
      const double win = std::pow(std::cos(t * PI / lpw / 2.0), 2.0);
      t *= PI;
      const double snc = (t == 0.0) ? 1.0 : std::sin(t) / t;
      K[(size_t)p * kw + j] = (float)(snc * win * (base / orig));
    }
  T->rs_kernel = e->upload(K);   // (previous table, if any, stays owned by the engine until destroy)
  T->rs_orig = orig; T->rs_new = nw; T->rs_width = width; T->rs_kw = kw;
}

__global__ void resample_kernel(const float* __restrict__ x, int N, const float* __restrict__ kern, int orig, int nw, int width, int kw,
                                float* __restrict__ out, int Lr) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Lr; i += gridDim.x * blockDim.x) {
    const int n = i / nw, p = i - n * nw;
    const float* kp = kern + (size_t)p * kw;
    const long base = (long)n * orig - width;
    float acc = 0.f;
    for (int j = 0; j < kw; ++j) {
      const long s = base + j;
      if (s >= 0 && s < N) acc = fmaf(kp[j], x[(size_t)b * N + s], acc);
    }
    out[(size_t)b * Lr + i] = acc;
  }
}
// windowed STFT frames with center=True reflect padding: frames[(b*F+f)][n] = w[n] * x[reflect(f*hop + n - 1024)]
__global__ void frames_kernel(const float* __restrict__ x, int L, const float* __restrict__ w, float* __restrict__ frames, int F) {
  const int bf = blockIdx.x, b = bf / F, f = bf % F;
  for (int n = threadIdx.x; n < MF_NFFT; n += blockDim.x) {
    long j = (long)f * MF_HOP + n - MF_NFFT / 2;
    if (j < 0) j = -j;
    if (j >= L) j = 2L * (L - 1) - j;
    frames[(size_t)bf * MF_NFFT + n] = w[n] * x[(size_t)b * L + j];
  }
}
__global__ void power_kernel(const float* __restrict__ spec, float* __restrict__ pw, long rows) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < rows * MF_NFREQ_PAD; i += (long)gridDim.x * blockDim.x) {
    const long r = i / MF_NFREQ_PAD;
    const int k = (int)(i - r * MF_NFREQ_PAD);
    float v = 0.f;
    if (k < MF_NFREQ) {
      const float re = spec[r * (2 * MF_NFREQ + 2) + k], im = spec[r * (2 * MF_NFREQ + 2) + MF_NFREQ + k];
      v = re * re + im * im;
    }
    pw[i] = v;
  }
}
// AmplitudeToDB('power', top_db=80): 10 log10(max(x,1e-10)), floored at (per-clip max - 80)
__global__ void __launch_bounds__(256) todb_kernel(float* __restrict__ mel, int F) {
  __shared__ float smax[8];
  const int b = blockIdx.x;
  float* p = mel + (size_t)b * F * MF_NMEL;
  const int n = F * MF_NMEL;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) {
    float v = 10.0f * log10f(fmaxf(p[i], 1e-10f));
    p[i] = v;
    mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = smax[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, smax[i]);
  const float floorv = mx - 80.0f;
  for (int i = threadIdx.x; i < n; i += 256) p[i] = fmaxf(p[i], floorv);
}

static void dense(ts_engine* e, const Layer& L, const float* A, float* C, long M, int ldc, cudaStream_t s) {
  GemmP p;
  p.A = A; p.W = L.W; p.C = C; p.M = (int)M; p.N = L.N; p.K = L.K; p.mper = (int)M;
  p.a_rs = L.K; p.kc = L.K; p.a_ts = L.K; p.c_rs = ldc; p.ldw = L.K;
  launch_gemm(e, p, s);
}

int mfcc_frames(int N, int sr0) {
  const int g = gcd_i(sr0, MF_SR), orig = sr0 / g, nw = MF_SR / g;
  const long Lr = ((long)nw * N + orig - 1) / orig;   // ceil(new * length / orig)
  return (int)(Lr / MF_HOP) + 1;
}

// wave [B,N] at sr0 -> Act3 [B, M, 64] (channel-last; pad rows as requested)
Act3 run_mfcc(ts_engine* e, const float* wave, int B, int N, int sr0, int pad, cudaStream_t s) {
  MfccTables* T = e->ws.sizing ? nullptr : tables(e);
  if (T) resample_kernel_for(e, T, sr0);
  const int g = gcd_i(sr0, MF_SR), orig = sr0 / g, nw = MF_SR / g;
  const int Lr = (int)(((long)nw * N + orig - 1) / orig);
  const int F = Lr / MF_HOP + 1;
  const long rows = (long)B * F;
  float* xr = e->ws.alloc<float>((size_t)B * Lr);
  float* frames = e->ws.alloc<float>((size_t)rows * MF_NFFT);
  float* spec = e->ws.alloc<float>((size_t)rows * (2 * MF_NFREQ + 2));
  float* pw = e->ws.alloc<float>((size_t)rows * MF_NFREQ_PAD);
  float* mel = e->ws.alloc<float>((size_t)rows * MF_NMEL);
  Act3 out = new_act(e, B, F, MF_NMFCC, pad, s);
  if (e->ws.sizing) return out;
  if (sr0 == MF_SR) {
    TS_CUDA(cudaMemcpyAsync(xr, wave, (size_t)B * N * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    resample_kernel<<<dim3(cdiv(Lr, 256), B), 256, 0, s>>>(wave, N, T->rs_kernel, orig, nw, T->rs_width, T->rs_kw, xr, Lr);
  }
  frames_kernel<<<(unsigned)rows, 256, 0, s>>>(xr, Lr, T->window, frames, F);
  e->launches += 2;
  dense(e, T->dft, frames, spec, rows, 2 * MF_NFREQ + 2, s);
  power_kernel<<<(int)std::min<long>((rows * MF_NFREQ_PAD + 255) / 256, 148 * 16), 256, 0, s>>>(spec, pw, rows);
  dense(e, T->mel, pw, mel, rows, MF_NMEL, s);
  todb_kernel<<<B, 256, 0, s>>>(mel, F);
  e->launches += 2;
  {  // DCT straight into the (padded) channel-last output
    GemmP p;
    p.A = mel; p.W = T->dct.W; p.C = out.row(0, 0); p.M = (int)rows; p.N = MF_NMFCC; p.K = MF_NMEL; p.mper = F;
    p.a_bs = (long)F * MF_NMEL; p.a_rs = MF_NMEL; p.kc = MF_NMEL; p.a_ts = MF_NMEL;
    p.c_bs = out.bstride(); p.c_rs = out.C; p.ldw = MF_NMEL;
    launch_gemm(e, p, s);
  }
  TS_CUDA(cudaGetLastError());
  return out;
}

}  // namespace ts

using namespace ts;

extern "C" int ts_mfcc_frames(int N, int sr) { return (N > 0 && sr > 0) ? mfcc_frames(N, sr) : 0; }

extern "C" int ts_mfcc(ts_engine* e, const float* wave, float* out, int B, int N, int sr, void* stream) {
  TS_API_BEGIN(e)
  if (e->host_only) fail(TS_ERR_UNSUPPORTED, "host-only engine cannot execute");
  if (B <= 0 || N < MF_NFFT || sr <= 0) fail(TS_ERR_INVALID, "ts_mfcc: B=%d N=%d sr=%d (need >= %d samples)", B, N, sr, MF_NFFT);
  cudaStream_t s = (cudaStream_t)stream;
  auto body = [&] {
    Act3 m = run_mfcc(e, wave, B, N, sr, 0, s);
    act_to_nct(e, m, MF_NMFCC, out, s);
  };
  e->ws.begin_sizing(); body();
  size_t need = e->ws.need;
  e->ws.buf.ensure(need + 256);
  e->ws.begin(need); body();
  TS_API_END(e)
}
