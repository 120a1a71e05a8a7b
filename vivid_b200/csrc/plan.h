// The recorded-plan object shared by plan.cu (replay, graphs, whole-call entry points) and netplan.cu (recording a plan from a
// network description inside the library).
#pragma once
#include <string>
#include <vector>

#include "common.h"

enum vb_op_kind { OP_CONV, OP_ATTN, OP_EW, OP_EMB, OP_PIN, OP_POUT, OP_HEUN };

struct vb_op {
  vb_op_kind kind;
  vb::ConvLaunch* conv = nullptr;
  union {
    vb_attn_desc attn;
    vb_ew_desc ew;
    vb_emb_desc emb;
    vb_precond_in_desc pin;
    vb_precond_out_desc pout;
    vb_heun_desc heun;
  };
  vb_op() { memset(&emb, 0, sizeof(emb)); }
};

struct vb_plan {
  std::vector<vb_op> ops;
  double flops = 0.0;
  int launches = 0;
  // one instantiated graph per replayed op range: the whole plan, and -- for no_time_enc feature caching
  // (generate_images.py:52-57) -- the source-view encoder and the denoising UNet on their own
  struct Range {
    int first, last;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
  };
  std::vector<Range> graphs;
  vb_io_desc io;
  bool io_bound = false;
  // plans recorded by the library itself (vb_net_plan_create) own their device buffers: prepared weights, activations, I/O
  std::vector<void*> owned;
  int enc_ops = 0;            // ops [0, enc_ops) are the source-view encoder
  // the encoder's cross-attention feature maps (16-bit NHWC [B][R][R][C]): what return_features / inject_features hand around
  struct Feature {
    void* ptr;
    int B, R, C;
  };
  std::vector<Feature> features;
  // how every prepared weight / converted buffer of the plan was produced, so that vb_net_plan_set_weights can refresh them in
  // place from another parameter table (the next checkpoint of the same architecture) without re-recording or re-tuning
  struct WeightSlot {
    std::string name;          // source parameter
    vb_weight_prep_desc d;     // dst and layout (src / src_dtype / gain are filled per refresh); copy: d.dst only
    std::string gain_param;    // 0-dim parameter the gain is multiplied with (emb_gain, out_gain), or empty
    double gain_const = 1.0;
    bool copy = false;         // plain conversion to fp32 (MPFourier buffers)
    long long numel = 0;
  };
  std::vector<WeightSlot> weights;
};
