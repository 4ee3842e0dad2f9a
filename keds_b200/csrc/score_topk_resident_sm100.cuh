// Scoring kernel for ONE query tile (<= 128 queries): the HBM-bound regime of the KEDs training
// step (128 queries vs 0.5M rows, src/trainer.py:213,221).
//
// The single-CTA kernel re-streams the 16-KB query k-block from L2 with every row tile (+50 %
// L2->SM traffic); at power-capped clocks that extra traffic costs 5-10 % of the HBM rate. Here a
// CTA PAIR (cluster of 2) issues one tcgen05.mma.cta_group::2 with M = 128: each CTA supplies 64
// query rows of operand A -- 96 KB of bf16 that stay RESIDENT in shared memory for the whole kernel
// -- and half (128 rows) of each 256-row tile of operand B. Per k-block only the 16-KB row half
// crosses L2->SM; nothing is read twice, from HBM or from L2.
//
// Accumulator layout of the M = 128 pair MMA per CTA ("2x2", 64 rows per CTA): TMEM lanes 0-63 hold
// (query m, tile rows 0..127) and lanes 64-127 hold (query m, tile rows 128..255), both in the same
// 128 columns. So an accumulator is 128 columns wide (4 fit in TMEM) and one query's scores of a
// tile are split between two epilogue threads (lane m and lane m + 64), each scanning its half of
// the tile. Each (query, half) keeps its own candidate list: a row slice of the planner becomes two
// "sub-slices" for the re-rank kernel, which only ever sees more slices.
//
//   warp 0      TMA producer (resident queries once, then the row half-tiles)
//   warp 1      MMA issuer (leader CTA only) + TMEM allocation (cta_group::2)
//   warps 2..5  epilogue
#pragma once
#include "score_topk_sm100.cuh"

namespace keds {

constexpr int RQ = 64;                         // queries per CTA
constexpr int R_STAGES = 4;
constexpr int R_NACC = 4;                      // accumulators of 128 columns
constexpr int R_ACC_COLS = BN / 2;             // 128
constexpr int R_CAP = 48;                      // candidate slots per (query, half)
constexpr int R_CHUNK = 16;                    // TMEM columns per tcgen05.ld
constexpr int R_MAX_KBLOCKS = 12;              // resident queries: 12 x 8 KB = 96 KB (d <= 768)
constexpr uint32_t R_QKB_BYTES = RQ * BK * 2;  // 8 KB: one query k-block of this CTA
constexpr uint32_t R_XBYTES = (BN / 2) * BK * 2;  // 16 KB: this CTA's half of a row tile k-block
constexpr uint32_t R_CAND_WARP_BYTES = R_CAP * 32 * 8;
constexpr uint32_t R_OFF_Q = 0;
constexpr uint32_t R_OFF_X = R_OFF_Q + R_MAX_KBLOCKS * R_QKB_BYTES;
constexpr uint32_t R_OFF_CAND = R_OFF_X + R_STAGES * R_XBYTES;
constexpr uint32_t R_OFF_BIAS = R_OFF_CAND + 4 * R_CAND_WARP_BYTES;
constexpr uint32_t R_OFF_BARS = R_OFF_BIAS + 2 * BN * 4;
constexpr uint32_t SCORE_RES_SMEM_BYTES = R_OFF_BARS + 256 + 1024;

// ScoreParams as for k_score_topk, with n_qt == n_qg == 1 and items = (db, slice). Candidate lines
// are written for sub-slice s' = 2 s + half: index ((db * 2S) + s') * 1 + 0.
__global__ void __launch_bounds__(SCORE_THREADS, 1)
k_score_topk_resident(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x0,
                      const __grid_constant__ CUtensorMap tm_x1, const ScoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));

  const uint32_t bars = sbase + R_OFF_BARS;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (R_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * R_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * R_STAGES + R_NACC + a); };
  const uint32_t qfull_bar = bars + 8u * (2 * R_STAGES + 2 * R_NACC);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + R_OFF_BARS + 200);
  volatile uint32_t* dead = reinterpret_cast<volatile uint32_t*>(gbase + R_OFF_BARS + 204);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();  // 0 = leader
  const int unit = static_cast<int>(blockIdx.x >> 1);
  const int n_units = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_q);
    prefetch_tensormap(&tm_x0);
    if (p.n_db > 1) prefetch_tensormap(&tm_x1);
    for (int s = 0; s < R_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < R_NACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // 4 epilogue warps of each CTA arrive on the leader's barrier
    }
    mbar_init(qfull_bar, 1);
    *dead = 0;
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  griddep_wait();               // bf16 queries come from k_prep_rows
  griddep_launch_dependents();
  const unsigned long long t_start = ktimer_begin(p.timing);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // resident operand A: this CTA's 64 queries, all k-blocks, once (bytes counted on the leader)
      if (crank == 0) mbar_arrive_expect_tx(qfull_bar, 2u * p.kblocks * R_QKB_BYTES);
      for (int kb = 0; kb < p.kblocks; ++kb)
        tma_load_2d_2sm(sbase + R_OFF_Q + kb * R_QKB_BYTES, &tm_q, qfull_bar, kb * BK,
                        static_cast<int>(crank) * RQ, kEvictLast);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = unit; item < p.n_items; item += n_units) {
        const ItemCoord c = decode_item(p, item);
        const CUtensorMap* tmx = c.db == 0 ? &tm_x0 : &tm_x1;
        for (int tile = c.t0; tile < c.t1; ++tile) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, dead, p.err, 0x100u + stage);
            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * R_XBYTES);
            tma_load_2d_2sm(sbase + R_OFF_X + stage * R_XBYTES, tmx, full_bar(stage), kb * BK,
                            tile * BN + static_cast<int>(crank) * (BN / 2), kEvictFirst);
            if (++stage == R_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(2 * RQ, BN);  // M = 128 over the pair, N = 256
      mbar_wait(qfull_bar, 0u, dead, p.err, 0x600u);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = unit; item < p.n_items; item += n_units) {
        const ItemCoord c = decode_item(p, item);
        for (int tile = c.t0; tile < c.t1; ++tile) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u, dead, p.err, 0x200u + acc);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * R_ACC_COLS);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(full_bar(stage), phase, dead, p.err, 0x300u + stage);
            tc_fence_after();
            const uint64_t adesc = smem_desc_sw128(sbase + R_OFF_Q + kb * R_QKB_BYTES);
            const uint64_t bdesc = smem_desc_sw128(sbase + R_OFF_X + stage * R_XBYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_bf16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(empty_bar(stage), 0x3);
            if (++stage == R_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_2sm(tfull_bar(acc), 0x3);
          if (++acc == R_NACC) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;                    // TMEM lane quadrant
    const int half = quad >> 1;                   // which 128-row half of the tile this thread sees
    const int q_local = static_cast<int>(crank) * RQ + (quad & 1) * 32 + lane;  // query in the tile
    const uint32_t wbuf = sbase + R_OFF_CAND + static_cast<uint32_t>(warp - 2) * R_CAND_WARP_BYTES;
    const uint32_t slot0 = wbuf + lane * 8;       // entry e of this lane lives at slot0 + e * 256
    float* sbias = reinterpret_cast<float*>(gbase + R_OFF_BIAS);
    const int et = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    int bias_buf = 0;
    for (int item = unit; item < p.n_items; item += n_units) {
      const ItemCoord c = decode_item(p, item);
      const bool active = q_local < p.nq;
      float theta = active ? -INFINITY : INFINITY;
      int cnt = 0;
      const float* bias = p.bias[c.db];
      const int n_rows = p.n_rows[c.db];
      for (int tile = c.t0; tile < c.t1; ++tile) {
        if (bias != nullptr) {
          sbias[bias_buf * BN + et] = bias[static_cast<long long>(tile) * BN + et];
          sbias[bias_buf * BN + 128 + et] = bias[static_cast<long long>(tile) * BN + 128 + et];
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        mbar_wait(tfull_bar(acc), acc_phase, dead, p.err, 0x400u + acc);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * R_ACC_COLS);
        const int row0 = tile * BN + half * (BN / 2);  // first DB row of this thread's columns
        const int nvalid = n_rows - row0;              // >= 128 when the half tile is full
#pragma unroll 1
        for (int ch = 0; ch < R_ACC_COLS / R_CHUNK; ++ch) {
          uint32_t v[R_CHUNK];
          tmem_ld16(taddr + ch * R_CHUNK, v);
          tmem_ld_wait16(v);
          if (bias != nullptr) {
            const float4* b4 =
                reinterpret_cast<const float4*>(sbias + bias_buf * BN + half * (BN / 2) + ch * R_CHUNK);
#pragma unroll
            for (int j4 = 0; j4 < R_CHUNK / 4; ++j4) {
              const float4 b = b4[j4];
              v[4 * j4 + 0] = __float_as_uint(__uint_as_float(v[4 * j4 + 0]) + b.x);
              v[4 * j4 + 1] = __float_as_uint(__uint_as_float(v[4 * j4 + 1]) + b.y);
              v[4 * j4 + 2] = __float_as_uint(__uint_as_float(v[4 * j4 + 2]) + b.z);
              v[4 * j4 + 3] = __float_as_uint(__uint_as_float(v[4 * j4 + 3]) + b.w);
            }
          } else if (nvalid < BN / 2) {
#pragma unroll
            for (int j = 0; j < R_CHUNK; ++j)
              if (ch * R_CHUNK + j >= nvalid) v[j] = 0xff800000u;  // -inf: zero-filled rows past the end
          }
          const uint32_t idx0 = static_cast<uint32_t>(row0 + ch * R_CHUNK);
          if (p.dump != nullptr && active) {
            float* drow = p.dump + (static_cast<long long>(c.db) * p.nq + q_local) * p.ld_dump;
#pragma unroll
            for (int j = 0; j < R_CHUNK; ++j)
              if (static_cast<int>(idx0) + j < n_rows) drow[idx0 + j] = __uint_as_float(v[j]);
          }
          uint32_t wptr = slot0 + static_cast<uint32_t>(cnt) * 256u;
#pragma unroll
          for (int j = 0; j < R_CHUNK; ++j) {
            if (__uint_as_float(v[j]) > theta) {
              sts64(wptr, v[j], idx0 + j);
              wptr += 256u;
            }
          }
          cnt = static_cast<int>((wptr - slot0) >> 8);
          if (__any_sync(0xffffffffu, cnt > R_CAP - R_CHUNK)) {
            const CandState st = compact_candidates<R_CAP>(slot0, cnt, theta);
            cnt = st.cnt;
            theta = st.theta;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
        if (++acc == R_NACC) {
          acc = 0;
          acc_phase ^= 1u;
        }
        bias_buf ^= 1;
      }
      if (__any_sync(0xffffffffu, cnt >= LKEEP)) {
        const CandState st = compact_candidates<R_CAP>(slot0, cnt, theta);
        cnt = st.cnt;
        theta = st.theta;
      }
      // sub-slice s' = 2 s + half of database db (one query tile): [(db * 2S + s')][query][LKEEP]
      const long long oitem = (static_cast<long long>(c.db) * (2 * p.S) + 2 * c.s + half);
      uint2* cbase = p.cand + (oitem * BM + q_local) * LKEEP;
#pragma unroll
      for (int e = 0; e < LKEEP; e += 2) {
        uint2 a = make_uint2(0xff800000u, 0xffffffffu), b = a;
        if (e < cnt) a = lds64(slot0 + e * 256);
        if (e + 1 < cnt) b = lds64(slot0 + (e + 1) * 256);
        *reinterpret_cast<uint4*>(cbase + e) = make_uint4(a.x, a.y, b.x, b.y);
      }
      p.cand_cnt[oitem * BM + q_local] = cnt;
      p.cand_theta[oitem * BM + q_local] = theta;
      __syncwarp();
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  ktimer_end(p.timing, t_start);
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace keds
