// Fused Gaussian-MLP policy + value forward for on-device rollout collection.
//
// Replaces, per env step of a rollout, RLlib's default fully connected model of the reference's low-level policy
// (REF train_config.py:91-113: fcnet_hiddens [256, 256], tanh, free log-std, vf_share_layers False) and the
// action sampling / log-probability code around it:
//
//     mean  = W3p tanh(W2p tanh(W1p obs + b1p) + b2p) + b3p          value = w3v . tanh(W2v tanh(W1v obs + b1v) + b2v) + b3v
//     a     = mean + exp(log_std) * noise          a_clipped = clip(a, -1, 1)          logp = sum_j N(a_j; mean_j, std_j)
//
// One CTA owns 128 envs (= the UMMA M dimension).  The six GEMMs run on the 5th-generation tensor cores: one thread
// issues tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation) with both operands in shared memory and the
// accumulators in tensor memory (layer 1 -> columns 0..255, layer 2 -> 256..511, layer 3 -> 0..31); the four warps
// read the accumulators back with tcgen05.ld, apply bias + tanh and write the next layer's A operand straight into
// shared memory in the canonical K-major core-matrix layout, so hidden activations never touch HBM.  Weights are
// pre-packed once (ilrl_policy_pack) into exactly that shared-memory image and streamed by cp.async.bulk through two
// 64 KB buffers, the next chunk in flight while the current layer's epilogue runs.
//
// Shared-memory operand layout (no swizzle, K-major): an R x K bf16 matrix is stored as [K/8][R][8] — 16-byte rows of
// 8 consecutive k, the R rows of one k-group contiguous.  In UMMA descriptor terms (cute/atom/mma_traits_sm100.hpp,
// "((8,n),2):((1,SBO),LBO)" in 16-byte units): SBO = 128 B between 8-row groups, LBO = 16 R bytes between the two
// k-halves of one K = 16 instruction; the next K step starts 32 R bytes further.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/ilrl.h"

namespace {

constexpr int TM = 128;    // envs per CTA
constexpr int NT = 256;    // threads per CTA: two per env row, each owning half of the 256 hidden columns
constexpr int HID = 256;   // hidden width
constexpr int K1 = 80;     // observation width padded to a multiple of 16 (70 low level, 44 high level)
constexpr int N3 = 32;     // output width padded (17 / 2 action means, 1 value)
constexpr uint32_t W1_BYTES = HID * K1 * 2, W2H_BYTES = 128 * HID * 2, W3_BYTES = N3 * HID * 2;
constexpr uint32_t NET_BYTES = W1_BYTES + 2 * W2H_BYTES + W3_BYTES;
constexpr uint32_t BIAS_WORDS = HID + HID + N3;
constexpr uint32_t BLOB_BIAS = 2 * NET_BYTES, BLOB_LOGSTD = BLOB_BIAS + 2 * BIAS_WORDS * 4, BLOB_BYTES = BLOB_LOGSTD + N3 * 4;

// shared memory map
constexpr uint32_t S_H = 0, S_A0 = S_H + TM * HID * 2, S_B0 = S_A0 + TM * K1 * 2, S_B1 = S_B0 + W2H_BYTES,
                   S_BIAS = S_B1 + W2H_BYTES, S_LOGSTD = S_BIAS + 2 * BIAS_WORDS * 4, S_BAR = S_LOGSTD + N3 * 4,
                   S_TMEM = S_BAR + 64, SMEM_BYTES = S_TMEM + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(W1_BYTES <= W2H_BYTES && W3_BYTES <= W2H_BYTES, "weight chunks must fit a buffer");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded: a protocol error must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 20)) {
      printf("ilrl_policy: mbarrier %u timed out (block %d thread %d)\n", bar, blockIdx.x, threadIdx.x);
      __trap();
    }
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  mbar_expect_tx(bar, bytes);
  bulk_copy(dst, src, bytes, bar);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle; rows = R of the operand image.  version 1 (Blackwell) in bits [46,48).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t rows) {
  const uint32_t lbo = rows * 16, sbo = 128;
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
  const uint32_t hi = (sbo >> 4) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
// D fp32, A/B bf16, both K-major, M = 128.
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((TM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[128 x n] (+)= A[:, 16 k0 .. 16 k1) * B[:, 16 k0 .. 16 k1)^T; A image has 128 rows, B image has n rows.  `fresh`: the
// first of these K steps overwrites D instead of accumulating.
__device__ __forceinline__ void gemm(uint32_t tmem_d, uint32_t a_s, uint32_t b_s, uint32_t n, int k0, int k1, bool fresh) {
  const uint32_t idesc = umma_idesc(n);
  const uint64_t ad = umma_desc(a_s, TM), bd = umma_desc(b_s, n);
  for (int k = k0; k < k1; ++k)   // descriptor address field is in 16-byte units: one K step = 32 * rows bytes
    umma(tmem_d, ad + (uint64_t)(k * 2 * TM), bd + (uint64_t)(k * 2 * n), idesc, !(fresh && k == k0));
}
// Issue only: the registers are valid after tmem_wait(v), which also ties them to the wait for the compiler.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));   // 2^-11 relative: below the bf16 rounding of the result
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// 32 accumulator columns of this thread's row -> tanh(x + bias) -> four 16-byte k-groups of the bf16 A operand image.
// The biases are read (8 x LDS.128, same address in every lane) BEFORE any store: a shared-memory load cannot be moved
// above a shared-memory store by the compiler, and interleaving them serialises load -> tanh -> store latencies.
__device__ __forceinline__ void hidden_chunk(const uint32_t (&v)[32], int c, const float* bias, uint8_t* H, int row) {
  float4 b[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) b[q] = *reinterpret_cast<const float4*>(bias + 32 * c + 4 * q);
  uint32_t p[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    p[2 * q] = pack_bf16(tanh_fast(__uint_as_float(v[4 * q]) + b[q].x), tanh_fast(__uint_as_float(v[4 * q + 1]) + b[q].y));
    p[2 * q + 1] = pack_bf16(tanh_fast(__uint_as_float(v[4 * q + 2]) + b[q].z), tanh_fast(__uint_as_float(v[4 * q + 3]) + b[q].w));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(H + ((4 * c + q) * TM + row) * 16) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
}
// accumulators [128 x 256] at `taddr` -> next layer's A operand H.  A thread owns one row and 128 columns = 4 chunks of
// 32; the TMEM load of the next chunk is in flight while the current one is processed.  After writing chunk i every
// thread arrives on mbarrier `bar_chunk + 8 i` and carries on; `consume(i)` (thread 0 waits for all 256 arrivals there)
// issues the next layer's K steps over exactly those columns — {2i, 2i+1} and {8+2i, 8+2i+1} — so that layer's
// tensor-core time hides under this epilogue and no warp but the issuing one ever waits for a slower warp.
template <class F>
__device__ __forceinline__ void epilogue_hidden(uint32_t taddr, const float* bias, uint8_t* H, int row, int half,
                                                uint32_t bar_chunk, F&& consume) {
  uint32_t va[32], vb[32];
  const int c0 = half * (HID / 64);
  tmem_ld32_issue(taddr + 32 * c0, va);
#pragma unroll
  for (int i = 0; i < HID / 64; i += 2) {
    tmem_wait(va);
    tmem_ld32_issue(taddr + 32 * (c0 + i + 1), vb);
    hidden_chunk(va, c0 + i, bias, H, row);
    fence_proxy_async();   // generic-proxy stores of this thread -> visible to the tensor core's async-proxy reads
    mbar_arrive(bar_chunk + 8 * i);
    consume(i);
    tmem_wait(vb);
    if (i + 2 < HID / 64) tmem_ld32_issue(taddr + 32 * (c0 + i + 2), va);
    hidden_chunk(vb, c0 + i + 1, bias, H, row);
    fence_proxy_async();
    mbar_arrive(bar_chunk + 8 * (i + 1));
    consume(i + 1);
  }
}

struct PolicyArgs {
  const uint8_t* blob;
  const float* obs;      // [n, obs_dim]
  const float* noise;    // [n, act_dim] or null (deterministic: a = mean)
  float* action;         // [n, act_dim] raw sample, or null
  float* action_clipped; // [n, act_dim] clip(a, -1, 1), or null
  float* logp;           // [n] or null
  float* value;          // [n] or null (skips the value net)
  int obs_dim, act_dim, n, nets;
  long long* timeline;   // harness: block 0 / thread 0 stamps clock64() at phase boundaries (null in normal use)
};

#define STAMP(i) do { if (A.timeline && blockIdx.x == 0 && tid == 0) A.timeline[i] = clock64(); } while (0)

__global__ void __launch_bounds__(NT, 1) policy_kernel(const PolicyArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (TM - 1), half = tid >> 7, env = blockIdx.x * TM + row;
  float* bias_s = reinterpret_cast<float*>(smem + S_BIAS);
  float* logstd_s = reinterpret_cast<float*>(smem + S_LOGSTD);
  const uint32_t bar_full0 = smem_u32(smem + S_BAR), bar_full1 = bar_full0 + 8, bar_mma = bar_full0 + 16, bar_in = bar_full0 + 24,
                 bar_chunk = bar_full0 + 32;   // four of them, one per 32-column chunk of a hidden epilogue
  const uint32_t sH = smem_u32(smem + S_H), sA0 = smem_u32(smem + S_A0), sB0 = smem_u32(smem + S_B0), sB1 = smem_u32(smem + S_B1);
  const int first = (A.nets & 1) ? 0 : 1, last = (A.nets & 2) ? 1 : 0;

  // A full, 16-byte aligned tile of observations is bulk-copied as raw fp32 into the (still unused) H region and
  // converted from there; a ragged last tile or an unaligned tensor is read with plain loads.
  STAMP(0);
  const int rows = min(TM, A.n - blockIdx.x * TM);
  const float* src = A.obs + (size_t)blockIdx.x * TM * A.obs_dim;
  const bool fast = rows == TM && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  constexpr uint32_t CONST_BYTES = (2 * BIAS_WORDS + N3) * 4;   // biases of both nets, then log-std: contiguous in both
  if (tid == 0) {
    mbar_init(bar_full0, 1); mbar_init(bar_full1, 1); mbar_init(bar_mma, 1); mbar_init(bar_in, 1);
    for (int i = 0; i < HID / 64; ++i) mbar_init(bar_chunk + 8 * i, NT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t tile_bytes = (uint32_t)(TM * A.obs_dim * 4);
    mbar_expect_tx(bar_in, CONST_BYTES + (fast ? tile_bytes : 0u));
    if (fast) bulk_copy(sH, src, tile_bytes, bar_in);
    bulk_copy(smem_u32(smem + S_BIAS), A.blob + BLOB_BIAS, CONST_BYTES, bar_in);
    bulk_load(sB0, A.blob + first * NET_BYTES, W1_BYTES, bar_full0);
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + S_TMEM)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (!fast)
    for (int i = tid; i < TM * K1 / 8; i += NT) *reinterpret_cast<uint4*>(smem + S_A0 + i * 16) = make_uint4(0, 0, 0, 0);
  __syncthreads();   // barrier inits visible to every thread; zero fill complete
  STAMP(1);
  mbar_wait(bar_in, 0);
  STAMP(2);
  // observation tile -> bf16 A operand image, zero padded in k (and for the rows past n)
  if (fast) {
    const float* rp = reinterpret_cast<const float*>(smem + S_H) + row * A.obs_dim;
#pragma unroll
    for (int kk = 0; kk < K1 / 16; ++kk) {
      const int kg = half * (K1 / 16) + kk;
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = kg * 8 + e < A.obs_dim ? rp[kg * 8 + e] : 0.f;
      *reinterpret_cast<uint4*>(smem + S_A0 + (kg * TM + row) * 16) =
          make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    }
  } else {
    for (int i = tid; i < rows * A.obs_dim; i += NT) {   // coalesced read of the contiguous tile
      const int r = i / A.obs_dim, k = i - r * A.obs_dim;
      *reinterpret_cast<__nv_bfloat16*>(smem + S_A0 + ((k >> 3) * TM + r) * 16 + (k & 7) * 2) = __float2bfloat16_rn(src[i]);
    }
  }
  fence_proxy_async();
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + S_TMEM);
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // a warp reaches TMEM lanes 32 (warp % 4) ..: its rows

  uint32_t ph0 = 0, ph1 = 0, phm = 0, phc = 0;   // ph0 / ph1 / phc are used by thread 0 only
  // (all CTAs start together: the 64 KB second-layer chunk is requested only now, behind the inputs of layer 1)
  if (tid == 0) bulk_load(sB1, A.blob + first * NET_BYTES + W1_BYTES, W2H_BYTES, bar_full1);
  __syncwarp();
  // noise tile: bulk-copied (full, aligned tiles) into the upper part of buffer 0, which is free whenever a W1 chunk
  // (40 KB of the 64) or nothing lives there; otherwise read with plain loads in the output epilogue
  const float* noise_src = A.noise ? A.noise + (size_t)blockIdx.x * TM * A.act_dim : nullptr;
  const bool noise_bulk = noise_src && rows == TM && (reinterpret_cast<uintptr_t>(noise_src) & 15) == 0;
  const float* noise_s = reinterpret_cast<const float*>(smem + S_B0 + W1_BYTES);
  float* stage_a = reinterpret_cast<float*>(smem + S_H);                 // output staging (H is free after layer 3)
  float* stage_c = reinterpret_cast<float*>(smem + S_H + TM * N3 * 4);

  STAMP(3);
  for (int net = first; net <= last; ++net) {
    const uint8_t* wb = A.blob + net * NET_BYTES;
    const float* bs = bias_s + net * BIAS_WORDS;
    const bool has_next = net < last;
    // layer 1: D[0..255] = obs * W1^T   (W1 in buffer 0)
    if (tid == 0) {
      mbar_wait(bar_full0, ph0); ph0 ^= 1;
      tc_after();
      STAMP(4 + 8 * net);
      gemm(tmem, sA0, sB0, HID, 0, K1 / 16, true);
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, phm); phm ^= 1;
    tc_after();
    STAMP(5 + 8 * net);
    if (tid == 0) bulk_load(sB0, wb + W1_BYTES + W2H_BYTES, W2H_BYTES, bar_full0);   // second half of W2
    __syncwarp();
    // layer 1 epilogue, with layer 2 issued slice by slice underneath it:
    // D[256..383] = H * W2[0..127]^T (buffer 1), D[384..511] = H * W2[128..255]^T (buffer 0)
    epilogue_hidden(tlane, bs, smem + S_H, row, half, bar_chunk, [&](int i) {
      if (tid == 0) {
        mbar_wait(bar_chunk + 8 * i, phc);
        tc_after();
        if (i == 0) { mbar_wait(bar_full1, ph1); ph1 ^= 1; mbar_wait(bar_full0, ph0); ph0 ^= 1; }
        gemm(tmem + 256, sH, sB1, 128, 2 * i, 2 * i + 2, i == 0);
        gemm(tmem + 256, sH, sB1, 128, 8 + 2 * i, 8 + 2 * i + 2, false);
        gemm(tmem + 384, sH, sB0, 128, 2 * i, 2 * i + 2, i == 0);
        gemm(tmem + 384, sH, sB0, 128, 8 + 2 * i, 8 + 2 * i + 2, false);
        if (i == HID / 64 - 1) umma_commit(bar_mma);
      }
      __syncwarp();
    });
    phc ^= 1;
    STAMP(6 + 8 * net);
    mbar_wait(bar_mma, phm); phm ^= 1;
    tc_after();
    STAMP(7 + 8 * net);
    if (tid == 0) {
      bulk_load(sB1, wb + W1_BYTES + 2 * W2H_BYTES, W3_BYTES, bar_full1);
      if (has_next) bulk_load(sB0, wb + NET_BYTES, W1_BYTES, bar_full0);
      if (net == 0 && noise_bulk) bulk_load(sB0 + W1_BYTES, noise_src, (uint32_t)(TM * A.act_dim * 4), bar_in);
    }
    __syncwarp();
    // layer 2 epilogue, with layer 3 underneath: D[0..31] = H * W3^T (buffer 1)
    epilogue_hidden(tlane + 256, bs + HID, smem + S_H, row, half, bar_chunk, [&](int i) {
      if (tid == 0) {
        mbar_wait(bar_chunk + 8 * i, phc);
        tc_after();
        if (i == 0) { mbar_wait(bar_full1, ph1); ph1 ^= 1; }
        gemm(tmem, sH, sB1, N3, 2 * i, 2 * i + 2, i == 0);
        gemm(tmem, sH, sB1, N3, 8 + 2 * i, 8 + 2 * i + 2, false);
        if (i == HID / 64 - 1) umma_commit(bar_mma);
      }
      __syncwarp();
    });
    phc ^= 1;
    STAMP(8 + 8 * net);
    mbar_wait(bar_mma, phm); phm ^= 1;
    tc_after();
    STAMP(9 + 8 * net);
    if (tid == 0 && has_next) bulk_load(sB1, wb + NET_BYTES + W1_BYTES, W2H_BYTES, bar_full1);
    __syncwarp();
    if (net == 0) STAMP(20);
    if (half == 0) {   // warp-uniform: warps 0..3 own the 128 rows of the narrow output layer
      uint32_t v[32];
      tmem_ld32_issue(tlane, v);
      if (net == 0 && noise_bulk) mbar_wait(bar_in, 1);
      if (net == 0) STAMP(21);
      tmem_wait(v);
      if (net == 0) STAMP(22);
      if (net == 0) {   // sample, clip, log-density; the action tiles are staged in shared memory (row stride act_dim)
        // branch-free over the padded width so that every load is issued before the first dependent use
        float mean[N3], z[N3], ls[N3];
#pragma unroll
        for (int j = 0; j < N3; ++j) {
          mean[j] = __uint_as_float(v[j]) + bs[2 * HID + j];
          ls[j] = logstd_s[j];
          z[j] = noise_bulk ? noise_s[row * A.act_dim + min(j, A.act_dim - 1)] : 0.f;
        }
        if (!noise_bulk && noise_src && env < A.n) {
#pragma unroll
          for (int j = 0; j < N3; ++j)
            if (j < A.act_dim) z[j] = noise_src[row * A.act_dim + j];
        }
        float lp = 0.f;
#pragma unroll
        for (int j = 0; j < N3; ++j) {
          const float a = fmaf(__expf(ls[j]), z[j], mean[j]);
          if (j < A.act_dim) {
            stage_a[row * A.act_dim + j] = a;
            stage_c[row * A.act_dim + j] = fminf(fmaxf(a, -1.f), 1.f);
            lp += -0.5f * z[j] * z[j] - ls[j] - 0.9189385332046727f;
          }
        }
        if (A.logp && env < A.n) A.logp[env] = lp;
      } else if (env < A.n) {
        A.value[env] = __uint_as_float(v[0]) + bs[2 * HID];
      }
    }
    if (net == 0) STAMP(23);
    tc_before();
    __syncthreads();   // D[0..31] is consumed (the next net's layer 1 overwrites it); the staged tiles are complete
    if (net == 0) STAMP(24);
    if (net == 0) {    // coalesced stores of the contiguous [rows, act_dim] tiles, loads batched ahead of the stores
      const int total = rows * A.act_dim;
      const size_t base = (size_t)blockIdx.x * TM * A.act_dim;
      constexpr int PER = TM * N3 / NT;
      float ta[PER], tc[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int i = min(tid + u * NT, TM * N3 - 1);
        ta[u] = stage_a[i];
        tc[u] = stage_c[i];
      }
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int i = tid + u * NT;
        if (A.action && i < total) A.action[base + i] = ta[u];
        if (A.action_clipped && i < total) A.action_clipped[base + i] = tc[u];
      }
      __syncthreads();   // before the next net's first epilogue writes H again
    }
    STAMP(10 + 8 * net);
  }
  if (warp == 0) {
    tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// fp32 torch-layout parameters ([out, in] row major) -> the blob of operand images, biases and log-std
__global__ void pack_kernel(const float* w1p, const float* b1p, const float* w2p, const float* b2p, const float* w3p,
                            const float* b3p, const float* w1v, const float* b1v, const float* w2v, const float* b2v,
                            const float* w3v, const float* b3v, const float* log_std, int obs_dim, int act_dim, uint8_t* blob) {
  const int net = blockIdx.y;
  const float* w1 = net ? w1v : w1p; const float* w2 = net ? w2v : w2p; const float* w3 = net ? w3v : w3p;
  const float* b1 = net ? b1v : b1p; const float* b2 = net ? b2v : b2p; const float* b3 = net ? b3v : b3p;
  const int out_dim = net ? 1 : act_dim;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(blob + net * NET_BYTES);
  const int total = (int)(NET_BYTES / 2);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // invert the image addressing: i -> (chunk, k-group, row, k within group)
    int e = i, rows, base_row = 0, which;
    if (e < (int)(W1_BYTES / 2)) { which = 0; rows = HID; }
    else if ((e -= W1_BYTES / 2) < (int)(W2H_BYTES / 2)) { which = 1; rows = 128; }
    else if ((e -= W2H_BYTES / 2) < (int)(W2H_BYTES / 2)) { which = 1; rows = 128; base_row = 128; }
    else { e -= W2H_BYTES / 2; which = 2; rows = N3; }
    const int kg = e / (rows * 8), r = (e / 8) % rows, k = kg * 8 + (e & 7);
    float x;
    if (which == 0) x = k < obs_dim ? w1[r * obs_dim + k] : 0.f;
    else if (which == 1) x = w2[(base_row + r) * HID + k];
    else x = r < out_dim ? w3[r * HID + k] : 0.f;
    img[i] = __float2bfloat16_rn(x);
  }
  float* bias = reinterpret_cast<float*>(blob + BLOB_BIAS) + net * BIAS_WORDS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)BIAS_WORDS; i += gridDim.x * blockDim.x)
    bias[i] = i < HID ? b1[i] : i < 2 * HID ? b2[i - HID] : (i - 2 * HID < out_dim ? b3[i - 2 * HID] : 0.f);
  if (net == 0 && blockIdx.x == 0 && threadIdx.x < N3)
    reinterpret_cast<float*>(blob + BLOB_LOGSTD)[threadIdx.x] = threadIdx.x < act_dim ? log_std[threadIdx.x] : 0.f;
}

long long* g_timeline = nullptr;

}  // namespace

extern "C" {

/* harness only, not in ilrl.h: device buffer of 32 int64 that block 0 of the next policy steps stamps with clock64() */
int ilrl_debug_policy_timeline(long long* dev32) { g_timeline = dev32; return ILRL_OK; }

int64_t ilrl_policy_blob_bytes(void) { return BLOB_BYTES; }

int ilrl_policy_pack(const float* w1_pi, const float* b1_pi, const float* w2_pi, const float* b2_pi, const float* w3_pi,
                     const float* b3_pi, const float* w1_vf, const float* b1_vf, const float* w2_vf, const float* b2_vf,
                     const float* w3_vf, const float* b3_vf, const float* log_std, int32_t obs_dim, int32_t act_dim,
                     void* blob_dev, void* stream) {
  if (!w1_pi || !b1_pi || !w2_pi || !b2_pi || !w3_pi || !b3_pi || !w1_vf || !b1_vf || !w2_vf || !b2_vf || !w3_vf ||
      !b3_vf || !log_std || !blob_dev || obs_dim < 1 || obs_dim > K1 || act_dim < 1 || act_dim > N3)
    return ILRL_ERR_ARG;
  pack_kernel<<<dim3(128, 2), 256, 0, (cudaStream_t)stream>>>(w1_pi, b1_pi, w2_pi, b2_pi, w3_pi, b3_pi, w1_vf, b1_vf, w2_vf,
                                                              b2_vf, w3_vf, b3_vf, log_std, obs_dim, act_dim, (uint8_t*)blob_dev);
  return cudaGetLastError() == cudaSuccess ? ILRL_OK : ILRL_ERR_CUDA;
}

int ilrl_policy_step(const void* blob_dev, const float* obs_dev, const float* noise_dev, float* action_dev,
                     float* action_clipped_dev, float* logp_dev, float* value_dev, int32_t obs_dim, int32_t act_dim,
                     int32_t n, void* stream) {
  const bool want_pi = action_dev || action_clipped_dev || logp_dev;
  if (!blob_dev || !obs_dev || n <= 0 || obs_dim < 1 || obs_dim > K1 || act_dim < 1 || act_dim > N3 ||
      (!want_pi && !value_dev) || ((uintptr_t)blob_dev & 15))
    return ILRL_ERR_ARG;
  static std::atomic<bool> configured[64];   // the opt-in to > 48 KB of dynamic shared memory is per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return ILRL_ERR_CUDA;
  if (!configured[dev].load(std::memory_order_acquire)) {   // idempotent: a concurrent first call just repeats it
    if (cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
      return ILRL_ERR_CUDA;
    configured[dev].store(true, std::memory_order_release);
  }
  PolicyArgs a;
  a.blob = (const uint8_t*)blob_dev; a.obs = obs_dev; a.noise = noise_dev; a.action = action_dev;
  a.action_clipped = action_clipped_dev; a.logp = logp_dev; a.value = value_dev;
  a.obs_dim = obs_dim; a.act_dim = act_dim; a.n = n; a.nets = (want_pi ? 1 : 0) | (value_dev ? 2 : 0); a.timeline = g_timeline;
  policy_kernel<<<(n + TM - 1) / TM, NT, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? ILRL_OK : ILRL_ERR_CUDA;
}

}  // extern "C"
