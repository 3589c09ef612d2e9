// Variable-length, non-causal, grouped-query attention (head_dim 64) on tcgen05 / TMEM / TMA,
// fused with the sigmoid output gate of Attn.forward.
//
// Replaces flash_attn_varlen_func(q, k, v, cu_seqlens, ...) * sigmoid(gate)
//   reference: model/base/transformer.py:100-103 (call), :87 (q|gate|k|v split)
//
// Layout: one packed activation buffer qkv[M, ld] bf16 with column blocks
//   [0,w) q (RoPE applied) | [w,2w) gate | [2w,2w+g) k (RoPE applied) | [2w+g,2w+2g) v
// Rows of one clip are contiguous (latent rows then patch rows); attention never crosses clips.
//
// One CTA = one 128-row query tile against the K/V stream of its clip, in 64-key sub-tiles; two CTAs per SM. The design
// (warp roles, the two softmax groups, the two softmax loops) is described above attn_fwd_kernel. A small pre-kernel
// (attn_kmax_kernel) provides max_j |k_j|^2 per (clip, kv head), from which every query row derives a bound on its scores.
#include "common.cuh"
#include "host_util.cuh"

namespace ttk {

struct AttnWork {
  int q_row0[2];   // first packed row of each query tile
  int q_valid[2];  // rows of the tile that belong to the clip (0 => tile unused)
  int q_head[2];   // query head of each tile
  int kv_head;
  int kv_row0;  // first packed row of the clip
  int kv_len;   // rows in the clip
  int kmax2;    // scratch, written by attn_kmax_kernel into the LEADER record of a (clip, kv head): float bits of
                //   max_j |k_j|^2 over the first half of the clip's keys ...
  int leader;   // index of that leader record (filled in by the planner)
  int kmax2b;   // ... and over the second half (two single-writer slots: two CTAs share the work, no atomics)
};
static_assert(sizeof(AttnWork) == 48, "AttnWork is mirrored in titok_video_b200/plan.py");

struct AttnParams {
  const AttnWork* work;
  const __nv_bfloat16* gate;  // qkv + width
  int64_t ld;                 // row pitch of qkv (elements)
  __nv_bfloat16* out;         // [M, ldo]
  int64_t ldo;
  float scale_log2;  // softmax_scale * log2(e)
  __nv_bfloat16* o_save;  // training: un-gated attention output [M, ldo] (null in inference)
  float* lse;             // training: [heads][M] log2-domain log-sum-exp of the scaled scores (null in inference)
  int M;
  const float* knorm2;  // optional [gqa / 64][M]: |k|^2 per row and kv head (written by the qkv GEMM epilogue); null: the
                        //   bound comes from attn_kmax_kernel through the leader work records
  long long* trace;  // development aid (ttk_debug_set_trace): [CTA][64] clock64 stamps of kv iterations 3..7
};

// slot = (j - 3) * 12 + event for kv iterations 3..7; events 0-5 softmax warp 4 (tile 0), 6-7 softmax warp 12 (tile 1),
// 8-11 MMA issuer
// (compiled in only with -DAT_TRACE: the stamps cost ~15 % in the softmax loop)
__device__ __forceinline__ void at_stamp(const AttnParams& p, int j, int ev) {
#ifdef AT_TRACE
  if (p.trace && j >= 3 && j < 7) p.trace[blockIdx.x * 64 + (j - 3) * 12 + ev] = clock64();
#endif
}

// CTA-level stamps (slots 48..): 48 kernel entry, 49 first sub-tile loop entered, 50 loop left, 51 epilogue stores issued
__device__ __forceinline__ void at_stamp_cta(const AttnParams& p, int slot) {
#ifdef AT_TRACE
  if (p.trace) p.trace[blockIdx.x * 64 + slot] = clock64();
#endif
}

constexpr int AT_BM = 128;  // query rows per tile
constexpr int AT_BN = 64;   // keys per kv sub-tile (one S accumulator)
constexpr int AT_D = 64;
constexpr int AT_KST = 6;  // K ring depth (sub-tiles)
constexpr int AT_VST = 4;  // V ring depth
constexpr int AT_Q_BYTES = AT_BM * AT_D * 2;   // 16 KB
constexpr int AT_KV_BYTES = AT_BN * AT_D * 2;  // 8 KB
constexpr int AT_XCH_BYTES = 4 * 128 * 4;  // per-row reference maximum, the two groups' row sums
// ~100 KB per CTA: two CTAs are resident per SM (2 x 256 tensor-memory columns, 2 x 384 threads)
constexpr int AT_SMEM = AT_Q_BYTES + (AT_KST + AT_VST) * AT_KV_BYTES + 512 + AT_XCH_BYTES + 1024;
constexpr int AT_THREADS = 128 + 8 * 32;
// TMEM columns (256 per CTA)
// group g (even / odd kv sub-tiles): S_g [96 g, 96 g + 64), P_g [96 g + 64, 96 g + 96) (64 keys x bf16 = 32 columns): P sits
// at the same distance behind S for both groups, so the softmax loop addresses both from one register
constexpr uint32_t AT_TM_S = 0;
constexpr uint32_t AT_TM_P = 64;
constexpr uint32_t AT_TM_GROUP = 96;
constexpr uint32_t AT_TM_O = 192;  // O [192,256)
constexpr uint32_t AT_TM_COLS = 256;
constexpr float AT_RESCALE_LOG2 = 8.0f;  // rescale O only when the row maximum grew by more than 2^8
// Bounded-score path (see attn_fwd_kernel): a query row whose Cauchy-Schwarz bound B = |q| max_j|k_j| scale log2(e) on its
// scaled scores is at most AT_BOUND_MAX uses the FIXED reference B - AT_BOUND_OFFSET for its exponentials:
// P = 2^(s c - B + OFFSET) <= 2^OFFSET can never overflow (row sums <= 2^(OFFSET+14), O <= 2^(OFFSET+14) |v|), and since
// the row maximum is >= -B the largest P of a row is >= 2^(OFFSET - 2 B) >= 2^-100: no underflow of the terms that matter.
constexpr float AT_BOUND_MAX = 90.0f;
constexpr float AT_BOUND_OFFSET = 80.0f;
// named barriers: token hand-over between the two softmax groups, end-of-loop exchange
constexpr uint32_t AT_BAR_TOK = 1;  // + group that WAITS on it (1, 2)
constexpr uint32_t AT_BAR_FIN = 3;

// (immediate barrier ids: a register id makes ptxas reserve all 16 named barriers of the CTA)
template <int ID>
__device__ __forceinline__ void nbar_sync256() {
  asm volatile("bar.sync %0, 256;" ::"n"(ID) : "memory");
}
template <int ID>
__device__ __forceinline__ void nbar_arrive256() {
  asm volatile("bar.arrive %0, 256;" ::"n"(ID) : "memory");
}

__device__ __forceinline__ int ld_global_s32(const int* ptr) {
  int v;
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// AND-reduction of a predicate over the 256 threads that meet at named barrier ID
template <int ID>
__device__ __forceinline__ bool bar_red_and_256(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n"
      ".reg .pred pi, po;\n"
      "setp.ne.u32 pi, %1, 0;\n"
      "bar.red.and.pred po, %2, 256, pi;\n"
      "selp.u32 %0, 1, 0, po;\n"
      "}\n"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(pred)), "n"(ID)
      : "memory");
  return out != 0;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef AT_EXP_NOMUFU  // timing experiment only (scripts/attn_bench.py): how much of the kernel is MUFU time?
  return fmaf(x, 1e-3f, 1e-2f);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// 2^x for a packed pair on the FMA / ALU pipes instead of MUFU (the kv loop is bound by the 16 ex2 / clk / SM of the XU
// pipe at head dim 64): x = n + f, n = rn(x) through the 1.5 * 2^23 magic constant, f in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5 -- P is rounded to bf16, 3.9e-3, right after),
// 2^n by adding n to the exponent field. x is clamped to >= -126 (masked keys hold -inf).
__device__ __forceinline__ void ex2_poly_pair(uint64_t t, float& o0, float& o1) {
  float x0, x1;
  f32x2_unpack(t, x0, x1);
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = f32x2_pack(x0, x1);
  const uint64_t u = f32x2_add(x, f32x2_pack(12582912.f, 12582912.f));    // low mantissa bits = rn(x)
  const uint64_t nf = f32x2_add(u, f32x2_pack(-12582912.f, -12582912.f));  // rn(x) as a float
  const uint64_t f = f32x2_fma(nf, f32x2_pack(-1.f, -1.f), x);
  uint64_t q = f32x2_fma(f32x2_pack(0.05517164617776871f, 0.05517164617776871f), f,
                         f32x2_pack(0.2426111251115799f, 0.2426111251115799f));
  q = f32x2_fma(q, f, f32x2_pack(0.6932609677314758f, 0.6932609677314758f));
  q = f32x2_fma(q, f, f32x2_pack(0.9999280571937561f, 0.9999280571937561f));
  float u0, u1, q0, q1;
  f32x2_unpack(u, u0, u1);
  f32x2_unpack(q, q0, q1);
  o0 = __uint_as_float((__float_as_uint(u0) << 23) + __float_as_uint(q0));
  o1 = __uint_as_float((__float_as_uint(u1) << 23) + __float_as_uint(q1));
}

// which of the 16 groups of four scores per thread take the polynomial: bit g of AT_EMU_B -> elements 4g+2, 4g+3,
// bit g of AT_EMU_A -> elements 4g, 4g+1. Default: 8 of 64 exponentials (12.5 %) off the MUFU pipe -- the loop is
// bound by instruction issue almost as much as by MUFU, and a polynomial pair costs ~14 issue slots against 5
// (measured at 64 clips, us per launch -- running-maximum loop: 0 % 366, 12.5 % 355, 25 % 371, 37.5 % 379, 50 % 405;
// bounded-score loop: 0 % 363, 12.5 % 332, 25 % 338, 37.5 % 336).
#ifndef AT_EMU_A
#define AT_EMU_A 0x0000
#endif
#ifndef AT_EMU_B
#define AT_EMU_B 0x1111
#endif

// Work item of a CTA: ONE 128-row query tile (tile blockIdx & 1 of work record blockIdx >> 1) against the K/V stream
// of its clip, in kv sub-tiles of 64 keys. Two CTAs are resident per SM, so one CTA's prologue (barrier init, tensor
// memory allocation, first loads) and epilogue (gate, normalisation, stores) run under the other's kv loop, and the
// two tiles of a record -- launched side by side -- find each other's K/V tiles in L2.
//
// Warp roles: warp 0 TMA producer (Q once, K and V sub-tiles through rings), warp 1 issues the score products, warp 3 the
// P V products (two in-order streams that do not wait for each other), warp 2 tensor-memory
// allocator, warps 4-7 softmax group a (even sub-tiles), warps 8-11 softmax group b (odd sub-tiles); thread == one
// query row x the 64 keys of a sub-tile, so a sub-tile's row maximum needs no exchange. Each group has its own S and P
// buffers in tensor memory (S(j+2) is issued as soon as S(j) sits in registers); both accumulate into the same O.
//
// (A persistent form -- one CTA per SM slot walking the work list, rings / S / P / O buffers running on across items, the
// next item's Q and first K / V sub-tiles fetched under the current item's epilogue -- was built and measured: 327 us per
// launch against 307 us at 64 clips. Two co-resident persistent CTAs fall into lock step and the per-item epilogue +
// prologue is not shorter than a fresh CTA's, whose set-up the other CTA of the SM covers just as well.)
//
// Two softmax loops, chosen per CTA before the first sub-tile:
//  * bounded-score loop (the common case). By Cauchy-Schwarz a row's scaled scores never exceed
//    B = |q| max_j |k_j| scale log2(e); with the FIXED per-row reference B - AT_BOUND_OFFSET the exponentials can neither
//    overflow nor (for B <= AT_BOUND_MAX) lose the terms that matter, so the loop needs no row maximum, no rescale of O
//    and no exchange between the groups: tensor-memory load, 64 exponentials, row sum, P back to tensor memory
//    (256 instead of 372 instructions per sub-tile and warp). Floating point makes the result independent of the
//    reference up to the rounding of s c - ref.
//  * running-maximum loop (rows with B > AT_BOUND_MAX, i.e. logits beyond +-60 nats), described next.
//
// Running-maximum loop: the exponential phases of the two groups are serialised by a token (named barriers): at any time at most one group
// of a CTA is on the MUFU pipe while the other one loads / reduces / stores, which keeps the pipe that bounds the loop
// busy instead of having all warps of a scheduler hit it in phase and then leave it idle together. The token also
// carries the per-row reference maximum (shared memory): the holder decides about the lazy rescale (only when a row
// maximum grew by more than 2^8), rescales O itself once every product issued so far has retired, and publishes the new
// reference before it hands the token over; the other group folds the change into its partial row sum when it next
// reads the reference. The final division by l makes the result exact.
template <bool TRAIN>  // TRAIN: also store the un-gated output and the log-sum-exp (compiled out of the inference kernel)
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  // (the work list is plan metadata, written long before the predecessor kernel: reading it ahead of pdl_wait is safe)
  const AttnWork* wp = p.work + (blockIdx.x >> 1);
  const int t = blockIdx.x & 1;
  const int q_valid = wp->q_valid[t];
  if (q_valid <= 0) return;  // (whole CTA: nothing has been allocated yet)
  const int q_row0 = wp->q_row0[t], q_head = wp->q_head[t];
  const int kv_head = wp->kv_head, kv_row0 = wp->kv_row0, kv_len = wp->kv_len;

  if (threadIdx.x == 128) at_stamp_cta(p, 48);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // [128][64]
  uint8_t* sK = sQ + AT_Q_BYTES;            // [KST][64][64]
  uint8_t* sV = sK + AT_KST * AT_KV_BYTES;  // [VST][64][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT_VST * AT_KV_BYTES);
  uint64_t* q_full = bars;              // [1]
  uint64_t* k_full = bars + 1;          // [KST]
  uint64_t* k_empty = k_full + AT_KST;  // [KST]
  uint64_t* v_full = k_empty + AT_KST;  // [VST]
  uint64_t* v_empty = v_full + AT_VST;  // [VST]
  uint64_t* s_full = v_empty + AT_VST;  // [2] S_g is in tensor memory
  uint64_t* s_empty = s_full + 2;       // [2] group g holds S_g in registers
  uint64_t* p_full = s_empty + 2;       // [2] P_g is in tensor memory
  uint64_t* pv_done = p_full + 2;       // [2] O += P_g V has retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* m_sh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [128] reference maximum per row
  float* l_sh = m_sh + 128;                                                        // [2][128] row sums of the groups

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_sub = (kv_len + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < AT_VST; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_empty[g], 4);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_done[g], 1);
    }
    fence_barrier_init();
    // The first loads wait for nothing inside the CTA: they go out before the tensor-memory allocation and the CTA-wide
    // barrier, so their L2 / HBM latency runs under the rest of the set-up. (Programmatic dependent launch: the
    // predecessor's results are needed from here.)
    pdl_wait();
    mbar_arrive_expect_tx(q_full, AT_Q_BYTES);
    tma_load_2d(sQ, &tmQ, q_full, q_head * AT_D, q_row0);
    for (int j = 0; j < AT_KST && j < n_sub; ++j) {
      mbar_arrive_expect_tx(&k_full[j], AT_KV_BYTES);
      tma_load_2d(sK + j * AT_KV_BYTES, &tmK, &k_full[j], kv_head * AT_D, kv_row0 + j * AT_BN);
      if (j < AT_VST) {
        mbar_arrive_expect_tx(&v_full[j], AT_KV_BYTES);
        tma_load_2d(sV + j * AT_KV_BYTES, &tmV, &v_full[j], kv_head * AT_D, kv_row0 + j * AT_BN);
      }
    }
  }
  if (warp == 2) tmem_alloc(tmem_ptr, AT_TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // programmatic dependent launch: the set-up above ran under the predecessor's tail

  // Register re-distribution: 384 threads start with 80 registers (two CTAs per SM); the control warpgroup drops to 32
  // and the 8 softmax warps (64 scores per thread live) grow to 104 (128 x 48 freed = 256 x 24 taken).
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (lane == 0) {  // (the thread that issued the first ring fill above)
      for (int j = AT_VST; j < n_sub; ++j) {
        const int sk = j % AT_KST, sv = j % AT_VST;
        if (j >= AT_KST) {
          mbar_wait(&k_empty[sk], ((j / AT_KST) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[sk], AT_KV_BYTES);
          tma_load_2d(sK + sk * AT_KV_BYTES, &tmK, &k_full[sk], kv_head * AT_D, kv_row0 + j * AT_BN);
        }
        mbar_wait(&v_empty[sv], ((j / AT_VST) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[sv], AT_KV_BYTES);
        tma_load_2d(sV + sv * AT_KV_BYTES, &tmV, &v_full[sv], kv_head * AT_D, kv_row0 + j * AT_BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);  // Q (K-major) x K (K-major)
      auto issue_s = [&](int g, int st) {
        const uint32_t sa = smem_u32(sQ);
        const uint32_t sb = smem_u32(sK + st * AT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16_ss(tmem_base + AT_TM_S + g * AT_TM_GROUP, umma_smem_desc_sw128(sa + k * 32, 1024, 0),
                       umma_smem_desc_sw128(sb + k * 32, 1024, 0), idesc_s, k != 0 ? 1u : 0u);
      };
      // Score products S(0), S(1), S(2), ... in order, each as soon as its K sub-tile has landed and the softmax group has
      // pulled the previous contents of its accumulator into registers. The P V products are issued by warp 3: a score
      // product never waits behind a P that the other softmax group hands over late (with one in-order stream
      // S(j+2), P(j) V(j) the softmax warps spent 11 % of their time waiting for scores).
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_sub; ++j) {
        const int g = j & 1, sk = j % AT_KST;
        mbar_wait(&k_full[sk], (j / AT_KST) & 1);
        if (j >= 2) mbar_wait(&s_empty[g], ((j - 2) >> 1) & 1);
        tc_fence_after();
        if (g == 0 && j >= 2) at_stamp(p, (j - 2) >> 1, 8);
        issue_s(g, sk);
        umma_commit(&s_full[g]);
        umma_commit(&k_empty[sk]);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (elect_one()) {
      // O += P(j) V(j), j ascending: the accumulation order into O is the same in every run and for every batch
      // composition, so results are bit-reproducible.
      constexpr uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_D, 0, 1);  // P (tmem, K-major) x V (MN-major)
      for (int j = 0; j < n_sub; ++j) {
        const int g = j & 1, sv = j % AT_VST;
        mbar_wait(&v_full[sv], (j / AT_VST) & 1);
        mbar_wait(&p_full[g], (j >> 1) & 1);
        tc_fence_after();
        if (g == 0) at_stamp(p, j >> 1, 10);
        const uint32_t sb = smem_u32(sV + sv * AT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)  // 16 keys == 8 packed columns of P
          umma_bf16_ts(tmem_base + AT_TM_O, tmem_base + AT_TM_P + g * AT_TM_GROUP + k * 8,
                       umma_smem_desc_sw128(sb + k * 2048, 1024, 0), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
        umma_commit(&pv_done[g]);
        umma_commit(&v_empty[sv]);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int g = (warp - 4) >> 2;      // softmax group: kv sub-tiles j == g (mod 2)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row within the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    // (pinned: left to itself the compiler re-derives these addresses from the thread index in every kv iteration -- the
    // loop is 250 instructions long and every one of them is worth 0.2 % of the launch)
    uint32_t t_s = tmem_base + lane_off + AT_TM_S + g * AT_TM_GROUP;
    const uint32_t t_o = tmem_base + lane_off + AT_TM_O;
    asm volatile("" : "+r"(t_s));
    const uint32_t t_p = t_s + (AT_TM_P - AT_TM_S);
    const float c = p.scale_log2;
    const uint64_t c2 = f32x2_pack(c, c);
    // shared-memory addresses of this group's barriers and of this row's exchange slots, converted once
    uint32_t a_sfull = smem_u32(&s_full[g]);
    asm volatile("" : "+r"(a_sfull));
    // the group's other barriers sit at fixed distances: s_full[2] | s_empty[2] | p_full[2] | pv_done[2] (8 bytes each)
    const uint32_t a_sempty = a_sfull + 16, a_pfull = a_sfull + 32, a_pvdone = a_sfull + 48;
    const uint32_t a_pvdone_o = a_pvdone + (g ? -8 : 8);
    const uint32_t a_msh = smem_u32(m_sh + r);
    // l_run: this group's share of the row sum, relative to m_ref (-inf until the group has seen a reference:
    // adopting one then scales the empty sum by 2^-inf = 0)
    float m_ref = __int_as_float(0xff800000), l_run = 0.f;

    // ---- which loop? The bound B_r = |q_r| max_j |k_j| c on this row's scaled scores (Cauchy-Schwarz; max_j |k_j|^2 of the
    // clip's kv head comes from attn_kmax_kernel through the leader work record). If every valid row of the tile has
    // B_r <= AT_BOUND_MAX the CTA takes the bounded-score loop: fixed per-row reference, no row maximum, no rescale, no
    // exchange between the groups. Otherwise (huge logits) it takes the running-maximum loop below. Both are exact up to
    // the bf16 rounding of P; the choice depends only on the clip's own q and k, never on what it is packed with.
    // The training forward (TRAIN) always takes the running-maximum loop: there the dominant probability of a row is
    // exactly 1, as in flash-attention, so its bf16 rounding pattern -- and with it the gradients of a chaotic network
    // (wide "stress" weights) -- follows the reference's (cosine 0.95..0.99 per parameter against 0.3..0.99 with the
    // bounded-score loop, whose dominant probability carries an ordinary 2^-9 rounding error; absolute accuracy of the two
    // loops against an fp64 softmax differs by 11 % at most, see profiles/r2_bench.md).
    float ref_l2 = 0.f;  // reference of the exponentials in log2 units (bounded-score loop)
    bool fast = false;
    if constexpr (!TRAIN) {
      // max_j |k_j|^2 (issued ahead of the wait for Q: the L2 latency of these loads runs under the arrival of the Q tile)
      float k2;
      if (p.knorm2) {
        // the qkv GEMM left |k|^2 of every row: the 256 softmax threads stride over the clip's rows
        const float* kn = p.knorm2 + static_cast<int64_t>(kv_head) * p.M + kv_row0;
        float mx = 0.f;
#pragma unroll 4
        for (int i = threadIdx.x - 128; i < kv_len; i += 256) mx = fmaxf(mx, __int_as_float(ld_global_s32(reinterpret_cast<const int*>(kn + i))));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) l_sh[warp - 4] = mx;
        nbar_sync256<AT_BAR_FIN>();
        const float4 m0 = *reinterpret_cast<const float4*>(l_sh), m1 = *reinterpret_cast<const float4*>(l_sh + 4);
        k2 = fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w)));
        // (l_sh is written again in the epilogue, behind the vote barrier below)
      } else {
        const AttnWork* lead = p.work + wp->leader;
        k2 = fmaxf(__int_as_float(ld_global_s32(&lead->kmax2)), __int_as_float(ld_global_s32(&lead->kmax2b)));
      }
      mbar_wait(q_full, 0);  // Q has landed in shared memory (async proxy -> mbarrier -> generic reads)
      float q2 = 0.f;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + r * 128 + ((cc ^ (r & 7)) << 4));
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = bf16_lo(w4[e]), x1 = bf16_hi(w4[e]);
          q2 = fmaf(x0, x0, q2);
          q2 = fmaf(x1, x1, q2);
        }
      }
      // 1.02: bf16 rounding of the RoPE-rotated q / k and the fp32 accumulation of the score never exceed it
      const float bound = sqrtf(q2 * k2) * c * 1.02f + 1e-3f;
      ref_l2 = bound - AT_BOUND_OFFSET;
      const bool row_ok = (r >= q_valid) || (bound <= AT_BOUND_MAX);  // (NaN / inf bounds fail the test)
#ifdef AT_NO_FAST  // (experiment / test hook: always take the running-maximum loop)
      fast = false;
      (void)row_ok;
#else
      fast = bar_red_and_256<AT_BAR_FIN>(row_ok);
#endif
    }
    // the epilogue's gate values: pull this thread's 64-byte segment into L2 ahead of its use
    {
      const __nv_bfloat16* gp = p.gate + static_cast<int64_t>(q_row0 + min(r, q_valid - 1)) * p.ld + q_head * AT_D + g * 32;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gp));
    }

    if (threadIdx.x == 128) at_stamp_cta(p, 49);
    if (fast) {
      // ===== bounded-score loop: P(j) = 2^(s c - ref), ref fixed per row =====
      const float nref = -ref_l2;
      const uint64_t nref2 = f32x2_pack(nref, nref);
      uint64_t la = f32x2_pack(0.f, 0.f), lb = la;
      const int last_valid = kv_len - (n_sub - 1) * AT_BN;  // keys of the last sub-tile (1..64)
      uint32_t par = 0;
      // exponentials of 32 scores sv[0..31] -> 16 packed bf16 pairs pk[0..15], row sums into la / lb
      auto exp32 = [&](uint32_t (&sv)[32], uint32_t (&pk)[16], const int half) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const uint64_t ta = f32x2_fma(f32x2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), c2, nref2);
          const uint64_t tb = f32x2_fma(f32x2_pack(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), c2, nref2);
          float a0, a1, b0, b1;
          if ((AT_EMU_A >> ((i >> 2) + 8 * half)) & 1) {
            ex2_poly_pair(ta, a0, a1);
          } else {
            f32x2_unpack(ta, a0, a1);
            a0 = ex2_approx(a0);
            a1 = ex2_approx(a1);
          }
          if ((AT_EMU_B >> ((i >> 2) + 8 * half)) & 1) {
            ex2_poly_pair(tb, b0, b1);
          } else {
            f32x2_unpack(tb, b0, b1);
            b0 = ex2_approx(b0);
            b1 = ex2_approx(b1);
          }
          la = f32x2_add(la, f32x2_pack(a0, a1));
          lb = f32x2_add(lb, f32x2_pack(b0, b1));
          pk[i >> 1] = pack_bf16x2(a0, a1);
          pk[(i >> 1) + 1] = pack_bf16x2(b0, b1);
        }
      };
      for (int j = g; j < n_sub; j += 2, par ^= 1) {
        // ---- S(j): this thread's 64 scores into registers, then hand S_g back to the tensor core
        mbar_wait_a(a_sfull, par);
        tc_fence_after();
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 0);
        uint32_t s0[32], s1[32], pk[32];
        tmem_ld_32x32b_x32(t_s, s0);
        tmem_ld_32x32b_x32(t_s + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_sempty);  // S(j+2) may overwrite the accumulator once the 4 warps arrived
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 1);
        if (j == n_sub - 1 && last_valid < 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= last_valid) s0[i] = 0xff800000u;  // -inf: exp2 -> 0
            if (32 + i >= last_valid) s1[i] = 0xff800000u;
          }
        }
        exp32(s0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]), 0);
        exp32(s1, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]), 1);
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 3);
        // ---- P(j) -> tensor memory once this group's previous P V no longer reads the buffer
        // (first pass: the wait for the phase before the barrier's first returns at once)
        mbar_wait_a(a_pvdone, par ^ 1);
        tc_fence_after();
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 4);
        // (deferring this hand-over to the top of the next iteration, so that the store's latency runs under the next
        // load, was measured: 334-341 vs 321 us per launch -- the P V product and everything queued behind it start later)
        tmem_st_32x32b_x32(t_p, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_pfull);
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 5);
      }
      {
        float x0, x1;
        f32x2_unpack(f32x2_add(la, lb), x0, x1);
        l_run = x0 + x1;
      }
    } else {
      for (int j = g; j < n_sub; j += 2) {
        const uint32_t par = (j >> 1) & 1;
        // ---- S(j): this thread's 64 scores into registers with one wait, then hand S_g back to the tensor core
        mbar_wait_a(a_sfull, par);
        tc_fence_after();
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 0);
        uint32_t sv[64];
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_sempty);  // S(j+2) may overwrite the accumulator once the 4 warps arrived
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 1);

        const int kv_valid = kv_len - j * AT_BN;  // < 64 only in the last sub-tile
        if (kv_valid < 64) {
  #pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kv_valid) sv[i] = 0xff800000u;  // -inf: exp2 -> 0, never the max
        }
        float m0 = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
        float m1 = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
  #pragma unroll
        for (int i = 4; i < 64; i += 4) {
          m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        }
        const float m_sub = fmaxf(m0, m1);  // a sub-tile always holds >= 1 valid key

        // ---- take the token: the other group's exponentials of sub-tile j-1 are done, its reference is published
        if (j == 0) {
          m_ref = m_sub;
          sts_f32(a_msh, m_sub);
        } else {
  #ifndef AT_EXP_NOTOKEN  // (timing experiment only: what does the token cost? results are wrong whenever a rescale fires)
          if (g == 0) nbar_sync256<AT_BAR_TOK>(); else nbar_sync256<AT_BAR_TOK + 1>();
  #endif
          const float m_cur = lds_f32(a_msh);
          if (m_cur != m_ref) {  // the other group moved the reference (rare)
            l_run *= ex2_approx((m_ref - m_cur) * c);
            m_ref = m_cur;
          }
          // lazy maximum: keep the reference unless some row of this warp outgrew it by 2^8
          // (rows past the clip's end hold another clip's queries: they must not take part in the decision, or the
          // rounding of this clip's rows would depend on what it is packed with)
          if (__any_sync(0xffffffffu, r < q_valid && (m_sub - m_ref) * c > AT_RESCALE_LOG2)) {
            const float m_new = fmaxf(m_ref, m_sub);
            const float alpha = ex2_approx((m_ref - m_new) * c);
            // O holds every product up to sub-tile j-1 once the latest PV of either group has retired
            mbar_wait_a(a_pvdone_o, ((j - 1) >> 1) & 1);
            if (j >= 2) mbar_wait_a(a_pvdone, ((j - 2) >> 1) & 1);
            tc_fence_after();
  #pragma unroll 1
            for (int q = 0; q < 4; ++q) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + q * 16, o);
              tmem_ld_wait();
  #pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x16(t_o + q * 16, o);
            }
            tmem_st_wait();
            tc_fence_before();
            l_run *= alpha;
            m_ref = m_new;
            sts_f32(a_msh, m_new);
          }
        }
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 2);

        // ---- P(j) = 2^(s*c - m*c): packed FMA, MUFU ex2 / polynomial, packed row-sum, bf16 pairs (in place, sv[0..31])
        const float nmc = -m_ref * c;
        const uint64_t nmc2 = f32x2_pack(nmc, nmc);
        uint64_t la = f32x2_pack(0.f, 0.f), lb = la;
  #pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const uint64_t ta = f32x2_fma(f32x2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), c2, nmc2);
          const uint64_t tb = f32x2_fma(f32x2_pack(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), c2, nmc2);
          float a0, a1, b0, b1;
          if ((AT_EMU_A >> (i >> 2)) & 1) {
            ex2_poly_pair(ta, a0, a1);
          } else {
            f32x2_unpack(ta, a0, a1);
            a0 = ex2_approx(a0);
            a1 = ex2_approx(a1);
          }
          if ((AT_EMU_B >> (i >> 2)) & 1) {
            ex2_poly_pair(tb, b0, b1);
          } else {
            f32x2_unpack(tb, b0, b1);
            b0 = ex2_approx(b0);
            b1 = ex2_approx(b1);
          }
          la = f32x2_add(la, f32x2_pack(a0, a1));
          lb = f32x2_add(lb, f32x2_pack(b0, b1));
          sv[i >> 1] = pack_bf16x2(a0, a1);
          sv[(i >> 1) + 1] = pack_bf16x2(b0, b1);
        }
        {
          float x0, x1;
          f32x2_unpack(f32x2_add(la, lb), x0, x1);
          l_run += x0 + x1;
        }
        // ---- hand the token over: the other group may start the exponentials of sub-tile j+1
  #ifndef AT_EXP_NOTOKEN
        if (j + 1 < n_sub) {
          if (g == 0) nbar_arrive256<AT_BAR_TOK + 1>(); else nbar_arrive256<AT_BAR_TOK>();
        }
  #endif
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 3);
        // ---- P(j) -> tensor memory once this group's previous P V no longer reads the buffer
        if (j >= 2) {
          mbar_wait_a(a_pvdone, ((j - 2) >> 1) & 1);
          tc_fence_after();
        }
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 4);
        tmem_st_32x32b_x32(t_p, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_pfull);
        if (threadIdx.x == 128) at_stamp(p, j >> 1, 5);
      }

    }

    if (threadIdx.x == 128) at_stamp_cta(p, 50);
    // ---- epilogue: out = bf16(O / l) * bf16(sigmoid(gate)); group g writes 32 of the 64 head dims
    pdl_launch_dependents();  // once the LAST CTAs of the grid are here, the successor may start setting itself up
    // this thread's gate values (prefetched into L2 before the loop): in flight while the groups meet at the barrier
    uint4 gv4[4];
    {
      const __nv_bfloat16* gt0 = p.gate + static_cast<int64_t>(q_row0 + min(r, q_valid - 1)) * p.ld + q_head * AT_D + g * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) gv4[q] = ldg16(gt0 + q * 8);
    }
    if (!fast) {  // (CTA-uniform)
      nbar_sync256<AT_BAR_FIN>();  // every exponential phase is over: the reference is final
      const float m_fin = lds_f32(a_msh);
      if (m_fin != m_ref) {
        l_run *= ex2_approx((m_ref - m_fin) * c);
        m_ref = m_fin;
      }
    }
    sts_f32(smem_u32(l_sh + g * 128 + r), l_run);
    nbar_sync256<AT_BAR_FIN>();
    const float l_tot = l_run + lds_f32(smem_u32(l_sh + (g ^ 1) * 128 + r));
    const float inv_l = __fdividef(1.0f, l_tot);
    if (TRAIN && g == 0 && r < q_valid)  // the backward kernels recompute P = 2^(s*c - lse)
      p.lse[static_cast<int64_t>(q_head) * p.M + q_row0 + r] = fast ? ref_l2 + __log2f(l_tot) : fmaf(m_ref, c, __log2f(l_tot));
    mbar_wait(&pv_done[(n_sub - 1) & 1], ((n_sub - 1) >> 1) & 1);  // the last product issued (in-order completion)
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32b_x32(t_o + g * 32, o);
    tmem_ld_wait();
    if (r < q_valid) {
      const int row = q_row0 + r;
      const int col = q_head * AT_D + g * 32;
      __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col;
      __nv_bfloat16* osv = TRAIN ? p.o_save + static_cast<int64_t>(row) * p.ldo + col : nullptr;
      // this thread's 32 head dims = two full 32-byte sectors of its row: 256-bit stores (a 128-bit store whose 32 lanes
      // hit 32 different rows costs the same 32 L1 wavefronts for half the bytes)
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        uint32_t ov[8], av[8];
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = 2 * q2 + qq;
          const uint4 gv = gv4[q];
          const uint32_t gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float g0 = bf16_lo(gg[e]), g1 = bf16_hi(gg[e]);
            // sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: one MUFU op per value instead of two (exp + reciprocal); the epilogue's
            // MUFU work was one extra kv sub-tile per row. tanh.approx is good to 2^-11, the result is rounded to bf16 (2^-9)
            const float s0 = bf16r(fmaf(0.5f, tanh_approx(0.5f * g0), 0.5f));
            const float s1 = bf16r(fmaf(0.5f, tanh_approx(0.5f * g1), 0.5f));
            const float a0 = bf16r(__uint_as_float(o[q * 8 + 2 * e]) * inv_l);
            const float a1 = bf16r(__uint_as_float(o[q * 8 + 2 * e + 1]) * inv_l);
            ov[qq * 4 + e] = pack_bf16x2(a0 * s0, a1 * s1);
            av[qq * 4 + e] = pack_bf16x2(a0, a1);
          }
        }
        stg32(dst + q2 * 16, ov);
        if (TRAIN) stg32(osv + q2 * 16, av);
      }
    }
  }

  if (threadIdx.x == 128) at_stamp_cta(p, 51);
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, AT_TM_COLS);
}

// max_j |k_j|^2 over the keys of one (clip, kv head), for the score bound of attn_fwd_kernel. A small persistent grid walks
// the work list; only the LEADER record of each (clip, kv head) -- the one whose own index the planner stored in `leader` --
// causes work: two CTAs take one half of the clip's keys each and keep their results (float bits) in the record's `kmax2`
// / `kmax2b` fields (single writers: no atomics, nothing to reset between launches). 8 lanes per key row (16 bytes each), 64 rows per pass. The maximum does not depend on the
// order of the rows: results are reproducible.
constexpr int AT_KMAX_THREADS = 512;
__global__ void __launch_bounds__(AT_KMAX_THREADS, 2)
attn_kmax_kernel(AttnWork* work, int n_work, const __nv_bfloat16* kbase, int64_t ld) {
  __shared__ float red[AT_KMAX_THREADS / 32];
  const int sub_row = (threadIdx.x & 31) >> 3;
  bool waited = false;
  for (int item = blockIdx.x; item < 2 * n_work; item += gridDim.x) {
    const int rec = item >> 1, half = item & 1;
    AttnWork* w = work + rec;
    // (the work list is plan metadata, written long before the predecessor kernel: reading it ahead of pdl_wait is safe)
    if (w->q_valid[0] <= 0 || w->leader != rec) continue;  // (CTA-uniform)
    if (!waited) {
      pdl_wait();  // the keys are the predecessor's output
      waited = true;
    }
    // this CTA's half of the clip's keys (the first half is rounded up to whole warps' rows)
    const int len = w->kv_len, split = ((len + 1) / 2 + 3) & ~3;
    const int row_begin = half ? split : 0;
    const int kv_len = (half ? len : (split < len ? split : len)) - row_begin;  // rows of this half (<= 0: nothing to do)
    const __nv_bfloat16* kp = kbase + static_cast<int64_t>(w->kv_row0 + row_begin) * ld + w->kv_head * AT_D + (threadIdx.x & 7) * 8;
    float mx = 0.f;
    constexpr int ROWS_PER_PASS = AT_KMAX_THREADS / 8, BATCH = 8;
    for (int base = (threadIdx.x >> 5) * 4; base < kv_len; base += BATCH * ROWS_PER_PASS) {  // (warp-uniform trip count)
      uint4 v[BATCH];  // all loads of a batch are in flight before the first use
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int row = base + u * ROWS_PER_PASS + sub_row;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (row < kv_len) v[u] = ldg16(kp + static_cast<int64_t>(row) * ld);
      }
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = bf16_lo(w4[e]), x1 = bf16_hi(w4[e]);
          ss = fmaf(x0, x0, ss);
          ss = fmaf(x1, x1, ss);
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        mx = fmaxf(mx, ss);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    __syncthreads();  // (red[] of the previous leader has been read)
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
      mx = threadIdx.x < AT_KMAX_THREADS / 32 ? red[threadIdx.x] : 0.f;
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (threadIdx.x == 0) (half ? w->kmax2b : w->kmax2) = __float_as_int(mx);
    }
  }
  // (a CTA that found no leader never waits: harmless, some other CTA of the grid did -- every valid record has a leader)
}

}  // namespace ttk

using namespace ttk;

extern "C" {

// qkv: packed [M, ld] bf16 (see header). work: device array of n_work AttnWork records (built by the
// host planner). out: [M, ldo] bf16 = attention(q,k,v) * sigmoid(gate).
static int attn_fwd_launch(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                           float softmax_scale, void* out, int64_t ldo, void* o_save, float* lse, const float* k_norm2,
                           cudaStream_t stream) {
  if (!qkv || !work || !out) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (width % 64 != 0 || gqa % 64 != 0 || ld % 8 != 0 || ldo % 16 != 0) return TTK_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(out) & 31u) != 0 || (reinterpret_cast<uintptr_t>(o_save) & 31u) != 0)
    return TTK_ERR_ALIGNMENT;  // 256-bit stores
  if (n_work <= 0) return TTK_OK;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  CUtensorMap tmQ, tmK, tmV;
  if (int e = make_tmap_bf16_2d(&tmQ, base, M, width, ld, AT_BM)) return e;
  if (int e = make_tmap_bf16_2d(&tmK, base + 2 * width, M, gqa, ld, AT_BN)) return e;  // boxes of 64 keys
  if (int e = make_tmap_bf16_2d(&tmV, base + 2 * width + gqa, M, gqa, ld, AT_BN)) return e;
  AttnParams p{};
  p.work = static_cast<const AttnWork*>(work);
  p.gate = base + width;
  p.ld = ld;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.trace = g_trace;
  p.o_save = static_cast<__nv_bfloat16*>(o_save);
  p.lse = lse;
  p.M = M;
  p.knorm2 = k_norm2;
  static PerDeviceOnce once_plain, once_train;
  if (int e = set_smem_attr_once(once_plain, reinterpret_cast<const void*>(attn_fwd_kernel<false>), AT_SMEM)) return e;
  if (int e = set_smem_attr_once(once_train, reinterpret_cast<const void*>(attn_fwd_kernel<true>), AT_SMEM)) return e;
  // score bound of every (clip, kv head): written into the scratch field of the leader work records (the work list is the
  // caller's device buffer; its `kmax2` fields are this library's scratch)
  if (!k_norm2 && !o_save) {  // (the training forward does not use the bound)
    const int cap = 2 * num_sms();
    if (int e = cuda_status(launch_pdl(attn_kmax_kernel, dim3(2 * n_work < cap ? 2 * n_work : cap), dim3(AT_KMAX_THREADS), 0, stream,
                                       const_cast<AttnWork*>(p.work), n_work, base + 2 * width, ld)))
      return e;
  }
  if (o_save)
    return cuda_status(launch_pdl(attn_fwd_kernel<true>, dim3(2 * n_work), dim3(AT_THREADS), AT_SMEM, stream, tmQ, tmK, tmV, p));
  return cuda_status(launch_pdl(attn_fwd_kernel<false>, dim3(2 * n_work), dim3(AT_THREADS), AT_SMEM, stream, tmQ, tmK, tmV, p));
}

int ttk_attn_varlen_fwd(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                        float softmax_scale, void* out, int64_t ldo, const float* k_norm2, cudaStream_t stream) {
  return attn_fwd_launch(qkv, ld, M, width, gqa, work, n_work, softmax_scale, out, ldo, nullptr, nullptr, k_norm2, stream);
}

// Training forward: additionally saves what the backward kernels (attn_bwd.cu) need: o_save [M, ldo] = the attention
// output before the gate, lse fp32 [width/64][M] = log2-domain log-sum-exp of the scaled scores.
int ttk_attn_varlen_fwd_train(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                              float softmax_scale, void* out, int64_t ldo, void* o_save, float* lse,
                              const float* k_norm2, cudaStream_t stream) {
  if (!o_save || !lse) return TTK_ERR_BAD_ARG;
  return attn_fwd_launch(qkv, ld, M, width, gqa, work, n_work, softmax_scale, out, ldo, o_save, lse, k_norm2, stream);
}

}  // extern "C"
