// distrifuser_b200 -- GroupNorm with cross-rank sufficient statistics (NHWC fp16).
// Replaces DistriGroupNorm.forward (distrifuser/modules/pp/groupnorm.py:14-97): the ~10 eager reduction /
// elementwise kernels and the 256-byte NCCL all_gather / all_reduce per layer become
//   (1) gn_stats_kernel    one HBM read, fp32 per-channel register accumulation, per-CTA partial moments; the LAST CTA to
//                          finish reduces the partials, exchanges (E[x], E[x^2]) with the patch group through peer stores +
//                          release/acquire flags over NVLink and applies the mode formula (gn_exchange)
//   (2) gn_apply_kernel    one read (L2-resident for <= ~60 MB activations) + one write, optional fused SiLU
// Both accept a per-(sample, channel) addend so that ResnetBlock2D's `conv1(x) + time_emb` never materialises.
#include <string.h>

#include "common.cuh"

using namespace df;

namespace {

#ifdef DF_GN_TRACE
// %globaltimer stamps of one launch for kernel tuning (tools/trace_gn.py); compiled out by default
__device__ unsigned long long df_gn_trace[16];
#define GN_TR(slot, cond) do { if (cond) df_gn_trace[slot] = globaltimer_ns(); } while (0)
#else
#define GN_TR(slot, cond) do {} while (0)
#endif

struct GnPlan {
  int V;        // 16-byte channel vectors per pixel (C/8)
  int lanes;    // pixels processed concurrently by one CTA
  int threads;  // blockDim
  int nchunk;   // CTAs per sample
  int ppc;      // pixels per CTA
};

inline GnPlan gn_plan(int b, int h, int w, int C) {
  GnPlan p;
  p.V = C / 8;
  p.lanes = 512 / p.V;
  if (p.lanes < 1) p.lanes = 1;
  int active = p.lanes * p.V;
  p.threads = (active + 31) / 32 * 32;
  int hw = h * w;
  int want = (2 * 148 + b - 1) / b;                  // ~2 CTAs per SM over the batch
  int cap = hw / (p.lanes * 8);                      // >= 8 pixels per thread
  if (cap < 1) cap = 1;
  p.nchunk = want < cap ? want : cap;
  p.ppc = (hw + p.nchunk - 1) / p.nchunk;
  p.nchunk = (hw + p.ppc - 1) / p.ppc;
  return p;
}

__device__ __forceinline__ void unpack8(const int4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

struct GnExchange {   // everything the last CTA needs to finish the statistics (was a separate 1-CTA kernel: 23 us of latency)
  df_comm_t c;
  float2* coef;
  unsigned int* ticket;
  int bG, nchunk_total;
  float inv_ne, bessel, eps;
  int mode, neg_fb, idx;
  uint64_t tensor_off, slot_bytes;
  uint32_t group_mask;
};

// Fused conv-halo handling of the normalise pass (GroupNorm -> SiLU -> 3x3 conv, the ResnetBlock2D pattern): the output goes
// into the interior rows of a [b, h+2, w, C] buffer, this rank's first / last output rows are ALSO stored into the patch
// neighbours' arena slots of the conv (replaces df_halo_push), and the two margin rows are filled from the neighbours' slots
// of the read epoch (replaces df_halo_assemble and its full-activation copy; distrifuser/modules/pp/conv2d.py:72-93).
struct GnHalo {
  int enabled;
  int h, w;                 // local rows, width
  int up, down;             // patch neighbours (communicator indices) or -1 at the image border
  int push;                 // ship this call's boundary rows (synchronous step: for this step; asynchronous: for the next one)
  int wait_flags;
  int idx;                  // comm tensor index of the CONV
  uint64_t off, slot_bytes;
  unsigned int* ticket2;    // CTA ticket of the normalise pass (self-resetting)
  df_comm_t c;
};

__device__ void gn_exchange(const GnExchange& e, const float2* __restrict__ partial, int G, int nchunk, float2* mine);

// Statistics pass of one CTA; returns true in the LAST CTA of the grid after it has run the exchange and written coef[].
// STREAM = true: loads bypass L1 (two-kernel path: the data is touched once); false: default caching, so that the apply pass
// of the fused kernel finds this CTA's pixels in L1 / L2.
template <bool STREAM>
__device__ __forceinline__ bool gn_stats_body(const __half* __restrict__ x, const __half* __restrict__ addend, int64_t addend_pitch,
                                              float2* __restrict__ partial, int hw, int C, int G, int V, int lanes,
                                              int ppc, const GnExchange& ex, float2* ch, bool partials_only = false) {
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x;
  const int v = tid % V, pl = tid / V;
  if (pl < lanes) {
    const int p0 = chunk * ppc, p1 = min(hw, p0 + ppc);
    const __half* base = x + ((size_t)b * hw) * C + (size_t)v * 8;
    float s[8], ss[8], ad[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = ad[j] = 0.f;
    if (addend) unpack8(ld_v4(addend + (size_t)b * addend_pitch + (size_t)v * 8), ad);   // per-(sample, channel) bias, e.g. the time embedding
    int p = p0 + pl;
    constexpr int U = 8;                          // loads in flight per thread (one DRAM latency round per 8 pixels)
    for (; p + (U - 1) * lanes < p1; p += U * lanes) {
      int4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = STREAM ? ld_nc_v4(base + (size_t)(p + u * lanes) * C) : ld_v4(base + (size_t)(p + u * lanes) * C);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[8];
        unpack8(r[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { float t = f[j] + ad[j]; s[j] += t; ss[j] = fmaf(t, t, ss[j]); }
      }
    }
    for (; p < p1; p += lanes) {
      float f[8];
      unpack8(STREAM ? ld_nc_v4(base + (size_t)p * C) : ld_v4(base + (size_t)p * C), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { float t = f[j] + ad[j]; s[j] += t; ss[j] = fmaf(t, t, ss[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ch[(size_t)pl * C + v * 8 + j] = make_float2(s[j], ss[j]);
  }
  GN_TR(1, blockIdx.x == 0 && blockIdx.y == 0 && tid == 0);
  __syncthreads();
  // fold channels x pixel-lanes into the G groups in a FIXED order (no atomics: results are bit-reproducible run to run)
  __shared__ float2 fold[256];
  const int cpg = C / G;
  const int parts = min(256 / G, 32);
  if (tid < parts * G) {
    const int g = tid % G, part = tid / G;
    const int per_group = cpg * lanes;
    float a = 0.f, q = 0.f;
    for (int e = part; e < per_group; e += parts) {
      const int pl2 = e / cpg, cc = e - pl2 * cpg;
      const float2 t = ch[(size_t)pl2 * C + g * cpg + cc];
      a += t.x; q += t.y;
    }
    fold[part * G + g] = make_float2(a, q);
  }
  __syncthreads();
  for (int g = tid; g < G; g += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int part = 0; part < parts; ++part) { a += fold[part * G + g].x; q += fold[part * G + g].y; }
    partial[((size_t)b * nchunk + chunk) * G + g] = make_float2(a, q);
  }
  if (partials_only) return false;                // fused kernel: grid barrier + per-CTA reduction follow (gn_fused_kernel)
  // last CTA of the grid finishes the job: reduce the partials, exchange with the patch group, write (mean, rstd)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int t = atomicAdd(ex.ticket, 1u);
    is_last = (t == (unsigned int)ex.nchunk_total - 1);
    if (is_last) *ex.ticket = 0;
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  gn_exchange(ex, partial, G, nchunk, ch);
  return true;
}

__global__ void __launch_bounds__(512) gn_stats_kernel(const __half* __restrict__ x, const __half* __restrict__ addend, int64_t addend_pitch,
                                                       float2* __restrict__ partial, int hw, int C, int G, int V, int lanes,
                                                       int ppc, GnExchange ex) {
  extern __shared__ float2 ch[];  // [lanes][C] per-channel (sum, sum of squares); reused as float2 mine[bG] by the last CTA
  gn_stats_body<true>(x, addend, addend_pitch, partial, hw, C, G, V, lanes, ppc, ex, ch);
}

// mode: 0 local, 1 synchronous exchange, 2 corrected_async_gn, 3 stale_gn   (see include/distrifuser_b200.h)
__device__ void gn_exchange(const GnExchange& e, const float2* __restrict__ partial, int G, int nchunk, float2* mine) {
  const df_comm_t& c = e.c;
  float2* __restrict__ coef = e.coef;
  const int bG = e.bG, mode = e.mode, neg_fb = e.neg_fb, idx = e.idx;
  const float inv_ne = e.inv_ne, bessel = e.bessel, eps = e.eps;
  const uint64_t tensor_off = e.tensor_off, slot_bytes = e.slot_bytes;
  const uint32_t group_mask = e.group_mask;
  const int tid = threadIdx.x, nthr = blockDim.x;
  // parallel reduction of the per-CTA partials: `tpp` threads per (sample, group) pair, then a shared-memory fold
  __shared__ float2 red[512];
  const int tpp = max(1, min(nthr / bG, 8));      // bG * tpp <= 512 (host guard: bG <= 512)
  for (int e = tid; e < bG * tpp; e += nthr) {    // stride loop: bG may exceed the block (b*groups = 512 on 320 threads)
    const int i = e / tpp, r = e - i * tpp;
    const int b = i / G, g = i - b * G;
    float s = 0.f, ss = 0.f;
    for (int k = r; k < nchunk; k += tpp) {
      float2 p = partial[((size_t)b * nchunk + k) * G + g];
      s += p.x; ss += p.y;
    }
    red[e] = make_float2(s, ss);
  }
  __syncthreads();
  for (int i = tid; i < bG; i += nthr) {
    float s = 0.f, ss = 0.f;
    for (int r = 0; r < tpp; ++r) { s += red[i * tpp + r].x; ss += red[i * tpp + r].y; }
    mine[i] = make_float2(s * inv_ne, ss * inv_ne);
  }
  __syncthreads();
  const int n = __popc(group_mask);
  uint32_t pub = 0, rd = 0;
  if (mode != 0) { pub = c.clock[0]; rd = c.clock[1]; }
  if (mode == 1) rd = pub;  // a synchronous exchange reads THIS epoch even inside an asynchronous step (sync_gn, groupnorm.py:74-80)

  auto publish = [&]() {
    for (int p = 0; p < c.world; ++p) {
      if (!(group_mask >> p & 1)) continue;
      float2* dst = (float2*)slot_ptr(c, p, pub, tensor_off, slot_bytes, c.rank);
      for (int i = tid; i < bG; i += nthr) dst[i] = mine[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < c.world && (group_mask >> tid & 1)) st_release_sys(c.flags[tid] + (size_t)idx * c.world + c.rank, pub);
  };
  auto wait_all = [&]() {
    if (tid < c.world && (group_mask >> tid & 1)) spin_until(c.flags[c.rank] + (size_t)idx * c.world + tid, rd, c.spin_timeout_ns);
    __syncthreads();
  };

  if (mode == 1) publish();          // fresh statistics are needed by everyone in this very step
  if (mode != 0) wait_all();         // sync: this epoch's; async: the previous epoch's (1-step stale)

  for (int i = tid; i < bG; i += nthr) {
    float2 m = mine[i];
    float mean = m.x, msq = m.y;
    if (mode != 0) {
      float sx = 0.f, sy = 0.f;
      float2 own_stale = make_float2(0.f, 0.f);
      for (int p = 0; p < c.world; ++p) {
        if (!(group_mask >> p & 1)) continue;
        float2 v = ((const float2*)slot_ptr(c, c.rank, rd, tensor_off, slot_bytes, p))[i];
        if (p == c.rank) own_stale = v;
        sx += v.x; sy += v.y;
      }
      const float invn = 1.f / (float)n;
      if (mode == 1) { mean = sx * invn; msq = sy * invn; }                                     // groupnorm.py:47,80
      else if (mode == 2) { mean = sx * invn + (m.x - own_stale.x); msq = sy * invn + (m.y - own_stale.y); }  // :49-51
      else { mean = (sx - own_stale.x + m.x) * invn; msq = (sy - own_stale.y + m.y) * invn; }  // :52-55
    }
    float var = msq - mean * mean;
    if (neg_fb && var < 0.f) var = m.y - m.x * m.x;                                             // :60-63
    var *= bessel;                                                                              // :65-66
    coef[i] = make_float2(mean, rsqrtf(var + eps));
  }
  if (mode >= 2) { __syncthreads(); publish(); }   // asynchronous: ship this step's statistics for the next step
}

// ---- fused kernel: statistics finished by EVERY CTA for its own sample (no serial exchange in one CTA, no coef[] round trip)
__device__ __forceinline__ float2 ld_cg_f2(const float2* p) {      // L2 load: the partials were written during this launch
  float2 r;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
  return r;
}
// (mean, mean of squares) of the G groups of sample b over the per-CTA partials, in a FIXED order (bit-reproducible):
// thread (g, r) sums the partials r, r + tpp, ... of group g (a warp reads 32 consecutive groups of one partial row: coalesced),
// then G threads fold the tpp pieces.  `red` holds G * tpp entries, `mine` G entries.
__device__ __forceinline__ void gn_reduce_sample(const float2* __restrict__ partial, int b, int G, int nchunk, float inv_ne,
                                                 float2* red, float2* mine) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tpp = max(1, min(nthr / G, 16));
  const float2* src = partial + (size_t)b * nchunk * G;
  for (int e = tid; e < G * tpp; e += nthr) {
    const int g = e % G, r = e / G;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    int k = r;
    for (; k + tpp < nchunk; k += 2 * tpp) {              // two loads in flight per thread and iteration
      const float2 u = ld_cg_f2(src + (size_t)k * G + g), w = ld_cg_f2(src + (size_t)(k + tpp) * G + g);
      s0 += u.x; q0 += u.y; s1 += w.x; q1 += w.y;
    }
    if (k < nchunk) { const float2 u = ld_cg_f2(src + (size_t)k * G + g); s0 += u.x; q0 += u.y; }
    red[r * G + g] = make_float2(s0 + s1, q0 + q1);
  }
  __syncthreads();
  for (int g = tid; g < G; g += nthr) {
    float s = 0.f, ss = 0.f;
    for (int r = 0; r < tpp; ++r) { s += red[r * G + g].x; ss += red[r * G + g].y; }
    mine[g] = make_float2(s * inv_ne, ss * inv_ne);
  }
  __syncthreads();
}

// This rank's statistics of ALL samples go to the patch group's slots of the publish epoch (one CTA of the grid does this:
// synchronous exchange -> at once, the peers are waiting; asynchronous modes -> for the next step, off the critical path).
__device__ void gn_publish_all(const GnExchange& e, const float2* __restrict__ partial, int G, int nchunk, float2* red, float2* mine_b) {
  const df_comm_t& c = e.c;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t pub = c.clock[0];
  const int nb = e.bG / G;
  for (int b = 0; b < nb; ++b) {
    gn_reduce_sample(partial, b, G, nchunk, e.inv_ne, red, mine_b);
    for (int p = 0; p < c.world; ++p) {
      if (!(e.group_mask >> p & 1)) continue;
      float2* dst = (float2*)slot_ptr(c, p, pub, e.tensor_off, e.slot_bytes, c.rank) + (size_t)b * G;
      for (int g = tid; g < G; g += nthr) dst[g] = mine_b[g];
    }
    __syncthreads();
  }
  __threadfence_system();
  __syncthreads();
  if (tid < c.world && (e.group_mask >> tid & 1)) st_release_sys(c.flags[tid] + (size_t)e.idx * c.world + c.rank, pub);
}

// (mean, rstd) of the G groups of sample b from this rank's statistics `mine` and, in the exchange modes, the patch group's
// slots of the read epoch (groupnorm.py:40-66; same arithmetic and order as gn_exchange).
__device__ __forceinline__ void gn_coef_sample(const GnExchange& e, int b, int G, const float2* mine, float2* coef_s) {
  const df_comm_t& c = e.c;
  const int tid = threadIdx.x, nthr = blockDim.x, mode = e.mode;
  uint32_t rd = 0;
  if (mode != 0) {
    rd = mode == 1 ? c.clock[0] : c.clock[1];   // a synchronous exchange reads THIS epoch even inside an asynchronous step
    if (tid < c.world && (e.group_mask >> tid & 1)) spin_until(c.flags[c.rank] + (size_t)e.idx * c.world + tid, rd, c.spin_timeout_ns);
    __syncthreads();
  }
  const int n = __popc(e.group_mask);
  for (int g = tid; g < G; g += nthr) {
    const float2 m = mine[g];
    float mean = m.x, msq = m.y;
    if (mode != 0) {
      float sx = 0.f, sy = 0.f;
      float2 own_stale = make_float2(0.f, 0.f);
      for (int p = 0; p < c.world; ++p) {
        if (!(e.group_mask >> p & 1)) continue;
        const float2 v = ((const float2*)slot_ptr(c, c.rank, rd, e.tensor_off, e.slot_bytes, p))[(size_t)b * G + g];
        if (p == c.rank) own_stale = v;
        sx += v.x; sy += v.y;
      }
      const float invn = 1.f / (float)n;
      if (mode == 1) { mean = sx * invn; msq = sy * invn; }
      else if (mode == 2) { mean = sx * invn + (m.x - own_stale.x); msq = sy * invn + (m.y - own_stale.y); }
      else { mean = (sx - own_stale.x + m.x) * invn; msq = (sy - own_stale.y + m.y) * invn; }
    }
    float var = msq - mean * mean;
    if (e.neg_fb && var < 0.f) var = m.y - m.x * m.x;
    var *= e.bessel;
    coef_s[g] = make_float2(mean, rsqrtf(var + e.eps));
  }
  __syncthreads();
}

template <bool STREAM>
__device__ __forceinline__ void gn_apply_body(const __half* __restrict__ x, const __half* __restrict__ addend, int64_t addend_pitch,
                                              __half* __restrict__ y, const __half* __restrict__ gamma,
                                              const __half* __restrict__ beta,
                                              const float2* __restrict__ coef, int hw, int C, int G, int V, int lanes,
                                              int ppc, int silu, const GnHalo& halo) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int v = tid % V, pl = tid / V;
  const size_t row_el = halo.enabled ? (size_t)halo.w * C : 0;          // elements of one image row
  uint32_t pub = 0;
  if (halo.enabled && halo.push) pub = halo.c.clock[0];
  if (pl < lanes) {
    const int cpg = C / G;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int ch = v * 8 + j;
      float2 mr;                       // generic load: coef[] is shared memory in the fused kernel, global otherwise
      asm volatile("ld.v2.f32 {%0, %1}, [%2];" : "=f"(mr.x), "=f"(mr.y) : "l"(coef + b * G + ch / cpg) : "memory");
      float ga = gamma ? __half2float(gamma[ch]) : 1.f, be = beta ? __half2float(beta[ch]) : 0.f;
      sc[j] = mr.y * ga;
      sh[j] = be - mr.x * sc[j];
    }
    if (addend) {                     // y = ((x + a) - mean) * rstd * gamma + beta  ==  x * sc + (sh + a * sc)
      float ad[8];
      unpack8(ld_v4(addend + (size_t)b * addend_pitch + (size_t)v * 8), ad);
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[j] = fmaf(ad[j], sc[j], sh[j]);
    }
    const int p0 = chunk * ppc, p1 = min(hw, p0 + ppc);
    const size_t base = ((size_t)b * hw) * C + (size_t)v * 8;
    // padded output: sample b starts at row b*(h+2), the interior at row 1
    const size_t ybase = halo.enabled ? ((size_t)b * (hw + 2 * halo.w) + halo.w) * C + (size_t)v * 8 : base;
    // neighbours' slots [2][batch][w*C]: part 0 = the sender's first row, part 1 = its last row (conv2d.py:61-65,90)
    __half* dst_up = nullptr;
    __half* dst_dn = nullptr;
    if (halo.enabled && halo.push) {
      const size_t nb = gridDim.y;
      if (halo.up >= 0) dst_up = (__half*)slot_ptr(halo.c, halo.up, pub, halo.off, halo.slot_bytes, halo.c.rank) + ((size_t)b) * row_el + (size_t)v * 8;
      if (halo.down >= 0) dst_dn = (__half*)slot_ptr(halo.c, halo.down, pub, halo.off, halo.slot_bytes, halo.c.rank) + (nb + b) * row_el + (size_t)v * 8;
    }
    const int last_row0 = hw - halo.w;                                  // first pixel of the last row (halo.enabled only)
    auto xform = [&](const int4& in) {
      float f[8];
      unpack8(in, f);
      int4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(f[j], sc[j], sh[j]);
        if (silu) t = __fdividef(t, 1.f + __expf(-t));
        f[j] = t;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
      return o;
    };
    auto emit = [&](int p, const int4& o) {
      st_v4(y + ybase + (size_t)p * C, o);
      if (dst_up && p < halo.w) st_v4(dst_up + (size_t)p * C, o);
      if (dst_dn && p >= last_row0) st_v4(dst_dn + (size_t)(p - last_row0) * C, o);
    };
    int p = p0 + pl;
    for (; p + 3 * lanes < p1; p += 4 * lanes) {
      int4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = STREAM ? ld_nc_v4(x + base + (size_t)(p + u * lanes) * C) : ld_v4(x + base + (size_t)(p + u * lanes) * C);
#pragma unroll
      for (int u = 0; u < 4; ++u) emit(p + u * lanes, xform(r[u]));
    }
    for (; p < p1; p += lanes) emit(p, xform(STREAM ? ld_nc_v4(x + base + (size_t)p * C) : ld_v4(x + base + (size_t)p * C)));
  }
  if (!halo.enabled) return;
  // ---- boundary rows shipped: the last CTA of the grid stamps the neighbours' flags (protocol of halo_push_kernel)
  if (halo.push) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned int tk = atomicAdd(halo.ticket2, 1u);
      if (tk == gridDim.x * gridDim.y - 1) {
        __threadfence();
        *halo.ticket2 = 0;
        if (halo.up >= 0) st_release_sys(halo.c.flags[halo.up] + (size_t)halo.idx * halo.c.world + halo.c.rank, pub);
        if (halo.down >= 0) st_release_sys(halo.c.flags[halo.down] + (size_t)halo.idx * halo.c.world + halo.c.rank, pub);
      }
    }
  }
  // ---- margin rows of sample b: first chunk fills the top one, last chunk the bottom one (zeros at the image border)
  const bool top = chunk == 0, bottom = chunk == (int)gridDim.x - 1;
  if (!top && !bottom) return;
  const uint32_t rd = (halo.up >= 0 || halo.down >= 0) ? halo.c.clock[1] : 0u;
  const size_t nb = gridDim.y;
  const int row_vec = (int)(row_el / 8);
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    if (side == 0 ? !top : !bottom) continue;
    const int src = side == 0 ? halo.up : halo.down;
    __half* dst = y + ((size_t)b * (halo.h + 2) + (side == 0 ? 0 : halo.h + 1)) * row_el;
    if (src >= 0) {
      if (halo.wait_flags && tid == 0) spin_until(halo.c.flags[halo.c.rank] + (size_t)halo.idx * halo.c.world + src, rd, halo.c.spin_timeout_ns);
      __syncthreads();
      // top margin = the up neighbour's LAST row (its part 1); bottom margin = the down neighbour's FIRST row (part 0)
      const __half* from = (const __half*)slot_ptr(halo.c, halo.c.rank, rd, halo.off, halo.slot_bytes, src) +
                           ((side == 0 ? nb : 0) + b) * row_el;
      for (int i = tid; i < row_vec; i += blockDim.x) st_v4(dst + (size_t)i * 8, ld_v4(from + (size_t)i * 8));
    } else {
      const int4 zero = make_int4(0, 0, 0, 0);
      for (int i = tid; i < row_vec; i += blockDim.x) st_v4(dst + (size_t)i * 8, zero);
    }
  }
}

__global__ void __launch_bounds__(512) gn_apply_kernel(const __half* __restrict__ x, const __half* __restrict__ addend, int64_t addend_pitch,
                                                       __half* __restrict__ y, const __half* __restrict__ gamma,
                                                       const __half* __restrict__ beta,
                                                       const float2* __restrict__ coef, int hw, int C, int G, int V, int lanes,
                                                       int ppc, int silu, GnHalo halo) {
  gn_apply_body<true>(x, addend, addend_pitch, y, gamma, beta, coef, hw, C, G, V, lanes, ppc, silu, halo);
}

// ONE launch: partial moments -> grid barrier -> every CTA finishes the statistics of ITS sample -> normalise.  Every CTA of the
// grid is resident at once (the host caps the grid at the occupancy of this kernel), so a CTA may wait on the generation word
// that the last arriver bumps.  After the barrier each CTA folds the per-CTA partials of its own sample itself (G x nchunk
// values from L2, a coalesced microsecond) and keeps (mean, rstd) in shared memory; the first design let the last CTA reduce
// everything, write coef[] and only then release the grid -- 6 us of serial work plus a global round trip in front of the
// normalise pass of all 296 CTAs (profiles/r2_gn_trace.txt).  Then each CTA normalises the pixels it has just read (L1 / L2
// hits: one HBM read and one write per element, one launch instead of two).  Exchange modes: the last arriver also publishes
// this rank's statistics to the patch group (synchronous mode: before anything else -- the peers wait for them).
__global__ void __launch_bounds__(512, 2) gn_fused_kernel(const __half* __restrict__ x, const __half* __restrict__ addend, int64_t addend_pitch,
                                                          __half* __restrict__ y, const __half* __restrict__ gamma,
                                                          const __half* __restrict__ beta, float2* __restrict__ partial,
                                                          int hw, int C, int G, int V,
                                                          int lanes, int ppc, int silu, GnExchange ex, unsigned int* gen,
                                                          GnHalo halo) {
  extern __shared__ float2 ch[];                         // [lanes][C] moments; after the fold: red[G*tpp] | mine[G] | coef[G]
  __shared__ unsigned int my_gen;
  __shared__ bool is_last;
  pdl_wait();
  GN_TR(0, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  if (threadIdx.x == 0) my_gen = ld_volatile_u32(gen);   // read before this CTA's ticket: the bump needs every CTA's ticket
  __syncthreads();
  gn_stats_body<false>(x, addend, addend_pitch, partial, hw, C, G, V, lanes, ppc, ex, ch, true);
  // ---- grid barrier: ticket, the last arriver resets it and bumps the generation
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ex.ticket, 1u);
    is_last = (t == (unsigned int)ex.nchunk_total - 1);
    if (is_last) {
      *ex.ticket = 0;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(gen), "r"(my_gen + 1u) : "memory");
    } else {
      unsigned int v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gen) : "memory");
        if (v == my_gen) __nanosleep(20);
      } while (v == my_gen);
    }
  }
  __syncthreads();
  GN_TR(2, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  const int tpp = max(1, min((int)blockDim.x / G, 16));
  float2* red = ch;
  float2* mine = ch + G * tpp;
  float2* coef_s = mine + G;
  const int nchunk = gridDim.x, b = blockIdx.y;
  if (is_last && ex.mode == 1) gn_publish_all(ex, partial, G, nchunk, red, mine);     // the peers (and this rank's CTAs) wait for it
  gn_reduce_sample(partial, b, G, nchunk, ex.inv_ne, red, mine);
  GN_TR(3, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  gn_coef_sample(ex, b, G, mine, coef_s);
  GN_TR(4, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  if (is_last && ex.mode >= 2) {                         // next step's statistics: only this CTA is late for its normalise pass
    float2* red2 = coef_s + G;                           // keep coef_s: scratch behind it (host sizes smem for both)
    gn_publish_all(ex, partial, G, nchunk, red2, red2 + G * tpp);
  }
  GN_TR(5, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  gn_apply_body<false>(x, addend, addend_pitch, y, gamma, beta, coef_s - (size_t)b * G, hw, C, G, V, lanes, ppc, silu, halo);
  GN_TR(6, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  GN_TR(7, is_last && threadIdx.x == 0);
}

#ifdef DF_GN_TRACE
}  // namespace
extern "C" int df_debug_gn_trace(unsigned long long* out_host /* 16 */) {
  DF_CHECK_CUDA(cudaDeviceSynchronize());
  DF_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, df_gn_trace, sizeof(unsigned long long) * 16));
  return 0;
}
namespace {
#endif

}  // namespace

extern "C" size_t df_groupnorm_scratch_bytes(int b, int groups, int h, int w, int C) {
  GnPlan p = gn_plan(b, h, w, C);
  return ((size_t)b * p.nchunk * groups + (size_t)b * groups) * sizeof(float2) + 256;
}

namespace {
int groupnorm_impl(df_comm_t comm, const void* x, const void* addend, int64_t addend_pitch, void* y, const void* gamma,
                   const void* beta, int b, int h, int w, int C, int groups, float eps, int mode, int bessel,
                   int neg_var_fallback, int fuse_silu, int idx, uint64_t tensor_off, uint64_t slot_bytes,
                   uint32_t group_mask, void* scratch, void* stream, GnHalo halo) {
  DF_REQUIRE(C % 8 == 0 && C % groups == 0 && C / 8 <= 512, "df_groupnorm_fwd: unsupported channel count %d", C);
  DF_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)addend % 16) == 0 && addend_pitch % 8 == 0,
             "df_groupnorm_fwd: x / y / addend must be 16-byte aligned (addend pitch a multiple of 8 elements)");
  if (addend == nullptr) addend_pitch = 0;
  else if (addend_pitch == 0) addend_pitch = C;
  DF_REQUIRE(mode >= 0 && mode <= 3, "df_groupnorm_fwd: bad mode %d", mode);
  DF_REQUIRE(b * groups <= 512 && groups <= 128, "df_groupnorm_fwd: b*groups = %d exceeds the exchange buffer", b * groups);
  DF_REQUIRE(mode == 0 || (slot_bytes >= (uint64_t)b * groups * 8 && (group_mask >> comm.rank & 1)),
             "df_groupnorm_fwd: statistics slot too small or rank outside its own group");
  cudaStream_t st = (cudaStream_t)stream;
  GnPlan p = gn_plan(b, h, w, C);
  // scratch: [ticket (256 B, zero-initialised by the caller once)] [partials] [coef]
  unsigned int* ticket = (unsigned int*)scratch;
  halo.ticket2 = ticket + 2;
  float2* partial = (float2*)((char*)scratch + 256);
  float2* coef = partial + (size_t)b * p.nchunk * groups;
  const int hw = h * w;
  const long long ne = (long long)(C / groups) * hw;
  GnExchange ex;
  ex.c = comm; ex.coef = coef; ex.ticket = ticket; ex.bG = b * groups; ex.nchunk_total = p.nchunk * b;
  ex.inv_ne = (float)(1.0 / (double)ne);
  ex.bessel = bessel ? (float)((double)ne / (double)(ne - 1)) : 1.f;
  ex.eps = eps; ex.mode = mode; ex.neg_fb = neg_var_fallback; ex.idx = idx;
  ex.tensor_off = tensor_off; ex.slot_bytes = slot_bytes; ex.group_mask = group_mask;
  size_t smem = (size_t)p.lanes * C * sizeof(float2);           // <= 32 KiB (lanes * C <= 4096)
  if (smem < (size_t)b * groups * sizeof(float2)) smem = (size_t)b * groups * sizeof(float2);
  {                                                             // fused kernel, after the fold: 2 x (red[G*tpp] + mine[G]) + coef[G]
    int tpp = p.threads / groups; if (tpp > 16) tpp = 16; if (tpp < 1) tpp = 1;
    const size_t need = (size_t)(2 * (groups * tpp + groups) + groups) * sizeof(float2);
    if (smem < need) smem = need;
  }
  // one launch when the whole grid is resident at once (always, for the plans of gn_plan on a B200: <= 2 CTAs per SM)
  static int fused_capacity = -1;
  if (fused_capacity < 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, 512, 32 * 1024) != cudaSuccess) per_sm = 0;
    fused_capacity = per_sm * sms;
    cudaGetLastError();
  }
  if (p.nchunk * b <= fused_capacity && smem <= 32 * 1024) {
    unsigned int* gen = ticket + 1;
    DF_CHECK_CUDA(launch_pdl(PDL_GN, gn_fused_kernel, dim3(p.nchunk, b), dim3(p.threads), smem, st, (const __half*)x, (const __half*)addend, addend_pitch,
                             (__half*)y, (const __half*)gamma, (const __half*)beta, partial, hw, C, groups, p.V,
                             p.lanes, p.ppc, fuse_silu, ex, gen, halo));
    return 0;
  }
  gn_stats_kernel<<<dim3(p.nchunk, b), p.threads, smem, st>>>((const __half*)x, (const __half*)addend, addend_pitch, partial, hw, C, groups,
                                                             p.V, p.lanes, p.ppc, ex);
  DF_CHECK_LAUNCH();
  gn_apply_kernel<<<dim3(p.nchunk, b), p.threads, 0, st>>>((const __half*)x, (const __half*)addend, addend_pitch, (__half*)y,
                                                           (const __half*)gamma, (const __half*)beta, coef, hw, C, groups, p.V,
                                                           p.lanes, p.ppc, fuse_silu, halo);
  DF_CHECK_LAUNCH();
  return 0;
}
}  // namespace

extern "C" int df_groupnorm_fwd(df_comm_t comm, const void* x, const void* addend, int64_t addend_pitch, void* y, const void* gamma,
                                const void* beta, int b, int h, int w, int C, int groups, float eps, int mode, int bessel,
                                int neg_var_fallback, int fuse_silu, int idx, uint64_t tensor_off, uint64_t slot_bytes,
                                uint32_t group_mask, void* scratch, void* stream) {
  GnHalo halo;
  memset(&halo, 0, sizeof(halo));
  return groupnorm_impl(comm, x, addend, addend_pitch, y, gamma, beta, b, h, w, C, groups, eps, mode, bessel, neg_var_fallback, fuse_silu, idx,
                        tensor_off, slot_bytes, group_mask, scratch, stream, halo);
}

extern "C" int df_groupnorm_halo_fwd(df_comm_t comm, const void* x, const void* addend, int64_t addend_pitch, void* y_padded, const void* gamma,
                                     const void* beta, int b, int h, int w, int C, int groups, float eps, int mode, int bessel,
                                     int neg_var_fallback, int fuse_silu, int idx, uint64_t tensor_off, uint64_t slot_bytes,
                                     uint32_t group_mask, void* scratch, int halo_idx, uint64_t halo_off,
                                     uint64_t halo_slot_bytes, int up_rank, int down_rank, int push, int wait_flags, void* stream) {
  DF_REQUIRE(halo_slot_bytes >= 2ull * b * w * C * 2, "df_groupnorm_halo_fwd: halo slot too small");
  DF_REQUIRE(up_rank < comm.world && down_rank < comm.world, "df_groupnorm_halo_fwd: neighbour outside the communicator");
  GnHalo halo;
  memset(&halo, 0, sizeof(halo));
  halo.enabled = 1; halo.h = h; halo.w = w; halo.up = up_rank; halo.down = down_rank; halo.push = push; halo.wait_flags = wait_flags;
  halo.idx = halo_idx; halo.off = halo_off; halo.slot_bytes = halo_slot_bytes; halo.c = comm;
  return groupnorm_impl(comm, x, addend, addend_pitch, y_padded, gamma, beta, b, h, w, C, groups, eps, mode, bessel, neg_var_fallback, fuse_silu,
                        idx, tensor_off, slot_bytes, group_mask, scratch, stream, halo);
}
