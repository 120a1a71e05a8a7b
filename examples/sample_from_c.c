/* From a checkpoint to samples through the C ABI alone (the same calls tests/test_netplan.py makes
 * through ctypes).  Compile check:  gcc -std=c11 -fsyntax-only -I include -I /usr/local/cuda/include examples/sample_from_c.c */
#include <stdio.h>
#include <cuda_runtime.h>
#include "vivid_b200.h"
/* params[]: every tensor of the checkpoint's state_dict under its reference name ("unet.enc.64x64_block0.conv_res0.weight",
 * "encoder.emb_fourier.freqs", ...), device pointers, fp32 or fp16.  net_desc: the constructor arguments (vivid-base:
 * img_resolution 64, model_channels 128, channel_mult {1,2,3,4}, num_blocks 3, attn_resolutions {16,8}, extra_attn 1, ...). */
int sample_batch(const vb_net_desc* base, const vb_param* base_params, int n_base,
                 const vb_net_desc* uncond, const vb_param* uncond_params, int n_uncond,
                 int batch, const float* src, const float* pose /*[batch,20]*/, const float* noise /*N(0,1) [batch,3,64,64]*/,
                 const float* t_steps /*HOST [33]*/, float* x_out, cudaStream_t stream, cudaStream_t side) {
  vb_plan *net = NULL, *gnet = NULL;
  int rc = vb_net_plan_create(base, base_params, n_base, batch, stream, &net);          /* once per (net, batch) */
  if (!rc) rc = vb_net_plan_create(uncond, uncond_params, n_uncond, batch, stream, &gnet);
  if (!rc) rc = vb_plan_set_inputs(net, src, pose, batch, NULL, stream);                /* constants of this sampler call */
  if (!rc) rc = vb_plan_set_inputs(gnet, NULL, NULL, 0, NULL, stream);                  /* gnet(src, x, t): no pose */
  float* ws = NULL;
  if (!rc) rc = cudaMalloc((void**)&ws, vb_sample_workspace_bytes(net)) ? VB_ERR_CUDA : 0;
  vb_sample_desc d = {0};
  d.net = net; d.gnet = gnet; d.noise = noise; d.t_steps = t_steps; d.workspace = ws; d.x_out = x_out;
  d.side_stream = side; d.num_steps = 32; d.guidance = 1.5f;
  if (!rc) rc = vb_sample(&d, stream);        /* 63 replays of each net + 63 vb_heun passes, enqueued; no host sync */
  if (rc) fprintf(stderr, "vivid_b200: %s\n", vb_last_error());
  /* ... cudaStreamSynchronize(stream) before x_out is read; vb_plan_destroy(net/gnet) frees the plans' device buffers */
  return rc;
}

/* The second stage (generate_images.py:318-331): the 64x64 sample is upscaled x4 (bilinear), the SR net denoises a fresh
 * 256x256 latent conditioned on it — drawing new N(0,1) conditioning noise before every call through `draw` (the reference
 * uses torch.randn_like, experiments/code/training/models.py:608-611; a C host plugs in its own generator) — and the result is
 * quantised to uint8 (training/encoders.py:58-62).  sr: a plan recorded from the SR checkpoint with super_res = 1. */
int super_resolve(vb_plan* sr, int batch, const float* base_sample /*[batch,3,64,64]*/, const float* sr_src /*[batch,3,256,256]*/,
                  const float* sr_pose, const float* sr_latent_noise, const float* t_steps, vb_noise_fn draw, void* draw_state,
                  float* cond /*scratch [batch,3,256,256]*/, float* ws, float* x256, uint8_t* images, cudaStream_t stream) {
  int rc = vb_resize(base_sample, cond, batch * 3, 64, 64, 256, 256, /*antialias=*/0, stream);
  if (!rc) rc = vb_plan_set_inputs(sr, sr_src, sr_pose, batch, cond, stream);
  vb_sample_desc d = {0};
  d.net = sr; d.noise = sr_latent_noise; d.t_steps = t_steps; d.workspace = ws; d.x_out = x256;
  d.sr_noise = draw; d.sr_noise_user = draw_state; d.num_steps = 32; d.guidance = 1.0f;      /* the SR stage is unguided */
  if (!rc) rc = vb_sample(&d, stream);
  if (!rc) rc = vb_decode_u8(x256, images, (int64_t)batch * 3 * 256 * 256, stream);
  return rc;
}
