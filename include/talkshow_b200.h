/* talkshow_b200 — C ABI of the B200-native TalkSHOW speech-to-motion engine.
 *
 * The reference (yhw-yhw/TalkSHOW) is pure Python/PyTorch and has no FFI of its own; the boundary
 * this library replaces is the *module level* of nets/spg/ — the calls the wrappers
 * nets/smplx_body_pixel.py, nets/smplx_body_vq.py and nets/smplx_face.py make.  Each entry point
 * below cites the reference call it replaces.  The host-side mirror of the wrappers lives in
 * talkshow_b200/nets/ (ctypes binding: talkshow_b200/_lib.py).
 *
 * Conventions
 *  - every function returns 0 (TS_OK) on success, a ts_status otherwise; ts_last_error() gives text;
 *  - all data pointers are DEVICE pointers on the engine's device unless the name says `host`;
 *    tensors are dense, fp32 / int64, in the reference's own layouts (channel-first [B,C,T]);
 *  - caller owns inputs/outputs; the engine owns packed weights + workspace;
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation;
 *  - one engine per device, not thread-safe.
 */
#ifndef TALKSHOW_B200_H
#define TALKSHOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ts_engine ts_engine;

enum ts_status {
  TS_OK = 0,
  TS_ERR_INVALID = 1,    /* bad argument / shape */
  TS_ERR_NOT_LOADED = 2, /* weights for this module were not loaded */
  TS_ERR_MISSING = 3,    /* a checkpoint tensor is missing or has the wrong shape */
  TS_ERR_CUDA = 4,       /* CUDA runtime error */
  TS_ERR_UNSUPPORTED = 5
};

/* One checkpoint tensor, HOST memory, fp32 (dtype 0) or int64 (dtype 1), dense row-major, named
 * exactly as in the reference's state_dict() after 'module.' stripping
 * (nets/smplx_body_pixel.py:117-126). */
typedef struct ts_tensor {
  const char* name;
  const void* data;
  int32_t dtype;
  int32_t ndim;
  int64_t shape[6];
} ts_tensor;

int ts_engine_create(ts_engine** out, int device);
void ts_engine_destroy(ts_engine* e);
const char* ts_last_error(ts_engine* e); /* e may be NULL: error of the last failed create */
int ts_engine_sm_count(ts_engine* e);

/* ---- weights: replace nn.Module.load_state_dict for each sub-module ------------------------ */
/* GatedPixelCNN(2048, dim, n_layers, 4, audio=True, bh_model=True): nets/smplx_body_pixel.py:53,
 * keys of ckpt['generator']['generator'].  dim/n_layers are read from the tensor shapes. */
int ts_load_pixelcnn(ts_engine* e, const ts_tensor* tensors, int n);
/* AudioEncoder(64,256,2,256): nets/smplx_body_pixel.py:46, ckpt['generator']['audioencoder']. */
int ts_load_audioenc(ts_engine* e, const ts_tensor* tensors, int n);
/* VQVAE(39|90,64,2048,1024,2,512): nets/smplx_body_pixel.py:54-62, which = 0 g_body, 1 g_hand. */
int ts_load_vq(ts_engine* e, int which, const ts_tensor* tensors, int n);
/* s2g_face.Generator: nets/smplx_face.py:37-45, ckpt['generator']['generator'].  The positional
 * conv must be passed with its effective weight under
 * 'audio_encoder.encoder.pos_conv_embed.conv.weight' [768,48,128] (the host shim resolves the
 * weight_norm parametrisation). */
int ts_load_face(ts_engine* e, const ts_tensor* tensors, int n);

/* ---- hot path ------------------------------------------------------------------------------- */
/* AudioEncoder.forward, nets/spg/vqvae_1d.py:27-34.  mfcc [B,64,M] -> out [B,256,T],
 * T = ts_latent_rows(M). */
int ts_audio_encode(ts_engine* e, const float* mfcc, float* out, int B, int M, void* stream);
int ts_latent_rows(int M);

/* GatedPixelCNN.generate, nets/spg/gated_pixelcnn_v2.py:152-177, as an exact O(T) incremental
 * evaluation.  aud [B,256,T0+T] (the AudioEncoder output, incl. the T0 prefix rows when
 * pre_latents is given, :158-165), label [B] int64, noise [2T,B,2048] = the Exp(1) draws the
 * reference's multinomial would consume, one [B,2048] block per sampled position in the order
 * (i,0),(i,1) (RNG contract), pre_latents [B,T0,2] int64 or NULL (T0=0).
 * idx_out [B,T,2] int64.  logits_out (may be NULL) [2T,B,2048]: the logits each draw used.
 * Device data is not range-checked: pre_latents (and `codes` of ts_pixelcnn_logits) must lie in [0,2048) like the
 * indices nn.Embedding accepts; labels outside [0, num_classes) are clamped. */
int ts_pixelcnn_generate(ts_engine* e, const float* aud, const int64_t* label, const float* noise,
                         int64_t* idx_out, float* logits_out, int B, int T, const int64_t* pre_latents, int T0,
                         void* stream);
/* GatedPixelCNN.forward (teacher forced), :130-150, for rows [0,T): logits_out [B,2048,T,2]. */
int ts_pixelcnn_logits(ts_engine* e, const float* aud, const int64_t* label, const int64_t* codes,
                       float* logits_out, int B, int T, void* stream);

/* VQVAE.decode(latents=...), nets/spg/vqvae_1d.py:201-208: idx [B,T] int64 -> out [B,C,4T]
 * (C = ts_vq_dim(which): 39 body / 90 hand, 78 / 180 for 6-D).  The indices are device data and are NOT range-checked (the
 * reference's F.embedding raises IndexError): every idx must lie in [0, num_embeddings) of the loaded codebook. */
int ts_vq_decode(ts_engine* e, int which, const int64_t* idx, float* out, int B, int T, void* stream);
/* VQVAE.encode, :196-199: poses [B,F,C] -> idx [B,T] int64 (T=F/4), e_out (may be NULL) [B,64,T]. */
int ts_vq_encode(ts_engine* e, int which, const float* poses, int64_t* idx, float* e_out, int B, int F,
                 void* stream);

/* s2g_face.Generator.forward, nets/spg/s2g_face.py:196-224: wave [B,N] (16 kHz), id [B,4] float
 * one-hot (zeros = "no id", nets/smplx_face.py:205-206) -> out [B,frame,103]. */
int ts_face_forward(ts_engine* e, const float* wave, const float* id, float* out, int B, int N, int frame,
                    void* stream);

/* pose channels of the loaded VQ-VAE `which` (in_dim of nets/smplx_body_pixel.py:54-57: 39 / 90 axis-angle,
 * 78 / 180 with convert_to_6d); 0 when not loaded. */
int ts_vq_dim(ts_engine* e, int which);

/* s2g_body_pixel.infer_on_audio core, nets/smplx_body_pixel.py:270-285, fused:
 * mfcc [B,64,M] -> codes [B,T,2] (may be NULL), poses [B,4T,C] with C = ts_vq_dim(0) + ts_vq_dim(1)
 * (body 39 + hand 90 = 129; 258 for the 6-D configs). */
int ts_body_generate(ts_engine* e, const float* mfcc, const int64_t* label, const float* noise, int64_t* codes,
                     float* poses, int B, int M, void* stream);

/* Audio front-end of get_mfcc_ta, data_utils/utils.py:148-177: wave [B,N] mono at
 * `sr` Hz -> torchaudio Resample(sr, 22000) -> MFCC(64 coefficients, n_fft 2048, hop 734 = 30 fps, 256 HTK
 * mels, top_db 80, DCT-II ortho) -> out [B,64,M], M = ts_mfcc_frames(N, sr). */
int ts_mfcc(ts_engine* e, const float* wave, float* out, int B, int N, int sr, void* stream);
int ts_mfcc_frames(int N, int sr);

/* scripts/demo.py:182-229 + data_utils/lower_body.py:68-87 (part2full): face [B,Ff,103],
 * body [B,Fb,129] -> out [B,Ff,265]; body is padded with its last frame / truncated to Ff. */
int ts_assemble_pose(ts_engine* e, const float* face, const float* body, float* out, int B, int Ff, int Fb,
                     int stand, void* stream);

/* scripts/demo.py:185-188,216-219 (convert_to_6d configs): matrix_to_axis_angle(rotation_6d_to_matrix(x)),
 * data_utils/rotation_conversion.py:512-533,433-447.  d6 [n,6] -> aa [n,3] (device pointers). */
int ts_rot6d_to_axis_angle(ts_engine* e, const float* d6, float* aa, int64_t n, void* stream);

/* ---- multi-GPU: the single collective of the path ------------------------------
 * The reference generates diversity samples / clips in a Python loop (scripts/demo.py:195-204); here they are sharded
 * one process per GPU and the [b,F,265] pose shards are all-gathered ONCE over NCCL (NVLink / NVSwitch).  NCCL is
 * dlopen'ed (libnccl_path: e.g. torch's nvidia/nccl/lib/libnccl.so.2; NULL = "libnccl.so.2" from the loader path).
 * Bootstrap: rank 0 calls ts_nccl_unique_id, ships the 128 bytes to the other ranks, every rank calls ts_nccl_init. */
int ts_nccl_unique_id(ts_engine* e, const char* libnccl_path, void* id128_host);
int ts_nccl_init(ts_engine* e, const char* libnccl_path, const void* id128_host, int rank, int world);
/* out[world*count] = concat over ranks of in[count] (fp32 device pointers), on `stream`. */
int ts_allgather(ts_engine* e, const float* in, float* out, int64_t count, void* stream);

/* ---- batched SMPL-X evaluation ---------------------------------------------------
 * Replaces the per-frame smplx_model(...) calls of scripts/demo.py:122-152 (get_vertices) and data_utils/get_j.py:20-51
 * (get_joints): smplx 0.1.28 SMPLX.forward + lbs (use_pca=False, flat_hand_mean=False, 300 betas, 100 expression
 * coefficients, static face landmarks) for F frames per call, fp32.  Model tensors (host, fp32 / int64), named as in
 * oracle/smplx_oracle.py: v_template [V,3], shapedirs [V,3,400], posedirs [486,3V], J_regressor [55,V], lbs_weights [V,55],
 * pose_mean [165], parents [55], faces [Fc,3], lmk_faces_idx [L], lmk_bary_coords [L,3], extra_joint_idx [E]. */
int ts_load_smplx(ts_engine* e, const ts_tensor* tensors, int n);
int ts_smplx_dims(ts_engine* e, int* V, int* njoints);
/* poses [F,265] in the reference's argument layout (jaw | leye | reye | global | body | lhand | rhand | expression,
 * demo.py:129-138), betas [300] or NULL (zeros, demo.py:159) -> vertices [F,V,3] (may be NULL), joints
 * [F,55+E+L,3] (may be NULL); use_expression = 0 evaluates with zero expression (get_vertices(exp=False)). */
int ts_smplx_forward(ts_engine* e, const float* poses, const float* betas, int use_expression, float* vertices,
                     float* joints, int F, void* stream);

/* ---- introspection (tests / bench) ---------------------------------------------------------- */
/* number of kernel launches issued by this engine since creation */
int64_t ts_launch_count(ts_engine* e);
/* device time of the last ts_pixelcnn_generate persistent-kernel launch is measured by the caller
 * with events; this returns the algorithmic weight bytes one latent row touches. */
int64_t ts_pixelcnn_row_bytes(ts_engine* e);
/* bytes the kernel actually stages per row (packed blob incl. row padding / per-column duplication) */
int64_t ts_pixelcnn_staged_row_bytes(ts_engine* e);
/* enable CUDA-event timing around the persistent kernel (events on its launch stream) and read the
 * duration of the most recent launch in ms (synchronises on the end event; -1 if none). */
int ts_pixelcnn_timing(ts_engine* e, int enable);
double ts_pixelcnn_last_ms(ts_engine* e);
/* Export the PixelCNN execution plan (stage table + packed weight blob) to host buffers so a test
 * can interpret it on the CPU; sizes are returned when the buffers are NULL. */
int ts_debug_pixelcnn_plan(ts_engine* e, int32_t* table, int64_t* table_len, float* blob, int64_t* blob_len);
/* dense C[M,N] = act(A[M,K] W[N,K]^T + bias) through one of the two GEMM kernels (unit tests):
 * mode 0 = fp32 FFMA kernel, 1 = tcgen05 3xTF32 kernel (K % 32 == 0), 2 = on-chip split (needs ts_set_tensor_cores(e, 5)),
 * 3 = tcgen05 fp16-split kernel (K % 64 == 0). */
int ts_debug_gemm(ts_engine* e, int mode, const float* A, const float* W, const float* bias, float* C, int M, int N,
                  int K, int act, void* stream);
/* Dense-contraction kernel for the face network and the VQ decoder (csrc/gemm_tc.cu):
 *   6 (default) tcgen05 CTA-pair kernel (cta_group::2, 256x256 tile) on two-term fp16-split operands (kind::f16, three
 *       products per MAC: fp32-grade results at twice the tf32 rate; operands must stay below 65504 in magnitude -- weights are
 *       pre-scaled per layer),
 *   1 / 3 the same kernel on 3xTF32 operands pre-split in HBM by the producing epilogue (full fp32 range),
 *   5 experiment: plain fp32 operands, tf32 hi / lo split by converter warps in shared memory (measured slower),
 *   4 single-CTA 128x256 kernel (tf32), 2 the same in clusters with TMA multicast, 0 everything on the fp32 FFMA2 kernel.
 * Every mode keeps the face outputs within 1e-4 of the reference (tests/test_gpu_parity.py). */
int ts_set_tensor_cores(ts_engine* e, int enable);
/* PixelCNN executor: 0 (default) = grid-wide persistent cooperative kernel (grid barrier; batch tile 16 / 32 / 64 picked per
 * launch), 1 = the same device code, one launch per stage (debug cross-check), 2 = cluster-resident executor: 16-CTA
 * clusters own 4 / 8 samples, TMA weight ring, stage hand-over through DSMEM mbarriers, any batch size and both
 * checkpoint geometries (dim 256 x 15 layers, dim 512 x 10 layers — the latter always runs here).  All three produce
 * bit-identical logits. */
int ts_set_pixelcnn_mode(ts_engine* e, int mode);
/* Debug: per-stage, per-CTA globaltimer stamps of one latent row of the persistent kernel.  ts_pixelcnn_trace(e, row)
 * arms it (row < 0 disarms) for the following ts_pixelcnn_generate calls; ts_pixelcnn_trace_read copies
 * out[stages][ctas][4] = {weights ready, left the grid barrier, task done, arrived} in ns (*len in/out, elements;
 * out may be NULL to query the size). */
/* stages per latent row and persistent CTAs of the loaded plan */
int ts_pixelcnn_plan_shape(ts_engine* e, int* nstages, int* ncta);
int ts_pixelcnn_trace(ts_engine* e, int row);
int ts_pixelcnn_trace_read(ts_engine* e, uint64_t* out, int64_t* len);

/* Persistent CTAs of the grid-wide plan built by the NEXT ts_load_pixelcnn (0 = one per SM).  A smaller even count (>= 64)
 * launches the sampler as CTA pairs on that many SMs and leaves the other TPCs free for kernels of another stream: with
 * the latency-bound sampler and the face regressor run side by side (talkshow_b200/pipeline.py: 64 clips 59.2 -> 51.0 ms). */
int ts_set_pixelcnn_ctas(ts_engine* e, int n);

/* ts_body_generate runs its two VQ-VAE decoders (body, hands — independent chains of ~20 small launches each) side by side
 * on two streams for batches of at most `max_batch` samples (default 16; 0 = one after the other on the caller's stream).  The
 * result is bit-identical either way (the chains share nothing but their input codes); the caller's stream is joined before the
 * call returns.  Measured: -0.6 ms per call at 1..16 samples, neutral at 32 / 64. */
int ts_set_vq_parallel(ts_engine* e, int max_batch);

/* Plan built by the NEXT ts_load_pixelcnn: 1 (default) = fused 52-stage plan (adjacent linear maps of the horizontal
 * stack multiplied together at load, layer-0 gate of column 1 gathered from a code table), 0 = plain 84-stage plan
 * (one stage per reference conv), 2 = EXPERIMENTAL: the fused plan with vert_to_horiz taken out of the vertical stages
 * (it runs one column at a time beside the horizontal stage before its consumer).  All evaluate
 * GatedPixelCNN.forward exactly up to fp32 rounding order. */
int ts_set_pixelcnn_fusion(ts_engine* e, int on);

#ifdef __cplusplus
}
#endif
#endif
