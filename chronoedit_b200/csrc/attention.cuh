// Interface of the tcgen05 attention forward (attention.cu).
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace ce {

struct AttnArgs {
  int B = 0, H = 0, Lq = 0, Lk = 0, head_dim = 128;
  const bf16* q = nullptr;  // q[b, i, h*hd + d] at q + (b*Lq + i)*ldq + h*hd + d
  int ldq = 0;
  const bf16* k = nullptr;  // k[b, j, h*hd + d] at k + (b*Lk + j)*ldk + ...
  int ldk = 0;
  const bf16* v = nullptr;
  int ldv = 0;
  // optional SECOND key/value source (Lk2 > 0): out = bf16(attn(q,k,v)) + bf16(attn(q,k2,v2)), two independent softmaxes
  // in ONE launch (text + image cross-attention, transformer_chronoedit.py:84-104)
  const bf16* k2 = nullptr;
  int ldk2 = 0;
  const bf16* v2 = nullptr;
  int ldv2 = 0;
  int Lk2 = 0;
  bf16* out = nullptr;      // out[b, i, h*hd + d]
  int ldo = 0;
  float scale = 0.f;        // 1/sqrt(head_dim)
  long long* timing = nullptr;  // optional [16] device counters: phase cycles of softmax warp 4 lane 0 of block (0,0,0) (profiling aid)
  // sequence-parallel output scatter (csrc/seqpar.cu): when peer_rows > 0, query row i of sample b is stored at
  //   out_peer[i / peer_rows] + ((b * peer_rows + i % peer_rows) * ldo + h * head_dim + out_col0)
  // i.e. straight into the attention-output buffer of the rank that owns that token (its own pointer or a peer's, NVLink).
  bf16* out_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int peer_rows = 0;
  int out_col0 = 0;
  int accumulate = 0;       // 1: out = bf16(float(bf16(attn)) + float(out))   (image + text cross-attention sum)
};

int launch_attention(const AttnArgs& a, cudaStream_t stream);
void set_attention_kernel(int version);  // -1 default / environment, 0, 2, 5, 6, 7 (see attention.cu)

}  // namespace ce
