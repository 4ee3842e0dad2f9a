// Scoring kernel for ONE query tile (<= 128 queries) with the query operand MULTICAST across a
// cluster of two CTAs: the HBM-bound regime of the KEDs training step (128 queries vs 0.5M rows,
// src/trainer.py:213,221).
//
// k_score_topk<false> re-reads the 16-KB query k-block from L2 with every 32-KB row k-block, i.e.
// one third of what L2 serves to the SMs is the same 192 KB over and over. Here two CTAs walk the
// SAME row slice (even / odd tiles) in lock step; per k-block each CTA fetches only HALF of the
// query box (64 queries, 8 KB) and the TMA unit multicasts it into both CTAs' shared memory. L2
// serves 40 KB per CTA and k-block instead of 48 KB; the MMAs stay single-CTA M = 128 x N = 256 at
// full tensor rate, and each CTA still keeps three 32-KB row boxes in flight.
//
// Hand-shake per pipeline stage (both CTAs run the same number of k-block steps):
//   full[s]   (this CTA)  1 arrival + 48 KB: own row box + both query halves (one from the peer)
//   empty[s]  (this CTA)  2 arrivals: own MMA commit and the peer's MMA commit (multicast commit) --
//             the peer's half lands in MY shared memory, so a stage is reusable only when both
//             CTAs' MMAs have consumed it
// A slice with an odd tile count gives one CTA a "dummy" step: it still exchanges query halves and
// recycles stages, but loads no rows and issues no MMA.
//
// Each CTA keeps its own candidate lists, so a planner slice turns into two sub-slices for the
// re-rank kernel: line index ((db * 2S) + 2s + cta) * n_qt + qt.
#pragma once
#include "score_topk_sm100.cuh"

namespace keds {

constexpr uint32_t QHALF_BYTES = (BM / 2) * BK * 2;  // 8 KB: 64 queries of one k-block

__global__ void __launch_bounds__(SCORE_THREADS, 1)
k_score_topk_mcast(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_x0,
                   const __grid_constant__ CUtensorMap tm_x1, const ScoreParams p) {
  using Cfg = ScoreCfg<false>;
  constexpr int NSTAGE = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));

  const uint32_t bars = sbase + Cfg::kOffBars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * NSTAGE + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * NSTAGE + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + Cfg::kOffBars + 104);
  volatile uint32_t* dead = reinterpret_cast<volatile uint32_t*>(gbase + Cfg::kOffBars + 108);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const int unit = static_cast<int>(blockIdx.x >> 1);
  const int n_units = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_q64);
    prefetch_tensormap(&tm_x0);
    if (p.n_db > 1) prefetch_tensormap(&tm_x1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    *dead = 0;
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers must exist before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  griddep_wait();  // bf16 queries come from k_prep_rows
  griddep_launch_dependents();
  const unsigned long long t_start = ktimer_begin(p.timing);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = unit; item < p.n_items; item += n_units) {
        const ItemCoord c = decode_item(p, item);
        const CUtensorMap* tmx = c.db == 0 ? &tm_x0 : &tm_x1;
        const int steps = (c.t1 - c.t0 + 1) >> 1;
        for (int i = 0; i < steps; ++i) {
          const int tile = c.t0 + 2 * i + crank;
          const bool real = tile < c.t1;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, dead, p.err, 0x100u + stage);
            const uint32_t sq = sbase + stage * Cfg::kStageBytes;
            mbar_arrive_expect_tx(full_bar(stage), real ? Cfg::kStageBytes : Q_STAGE_BYTES);
            tma_load_2d_mcast(sq + crank * QHALF_BYTES, &tm_q64, full_bar(stage), kb * BK,
                              c.qg * BM + crank * (BM / 2), 0x3, kEvictLast);
            if (real)
              tma_load_2d(sq + Q_STAGE_BYTES, tmx, full_bar(stage), kb * BK, tile * BN, kEvictFirst);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
      // tail: every stage released by both CTAs, i.e. no commit of the peer is still on its way
      // to this CTA's barriers when the cluster leaves
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_wait(empty_bar(stage), phase ^ 1u, dead, p.err, 0x500u + stage);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = unit; item < p.n_items; item += n_units) {
        const ItemCoord c = decode_item(p, item);
        const int steps = (c.t1 - c.t0 + 1) >> 1;
        for (int i = 0; i < steps; ++i) {
          const bool real = c.t0 + 2 * i + crank < c.t1;
          uint32_t d_tmem = 0;
          if (real) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u, dead, p.err, 0x200u + acc);
            tc_fence_after();
            d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
          }
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(full_bar(stage), phase, dead, p.err, 0x300u + stage);
            tc_fence_after();
            if (real) {
              const uint32_t sq = sbase + stage * Cfg::kStageBytes;
              const uint64_t adesc = smem_desc_sw128(sq);
              const uint64_t bdesc = smem_desc_sw128(sq + Q_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UK; ++k)
                umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_mcast(empty_bar(stage), 0x3);  // frees the stage in both CTAs
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1u;
            }
          }
          if (real) {
            umma_commit(tfull_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (lane == query)
    const int quad = warp & 3;
    const int q_local = quad * 32 + lane;
    const uint32_t wbuf = sbase + Cfg::kOffCand + static_cast<uint32_t>(warp - 2) * CAND_WARP_BYTES;
    const uint32_t slot0 = wbuf + lane * 8;  // entry e of this lane lives at slot0 + e * 256
    float* sbias = reinterpret_cast<float*>(gbase + Cfg::kOffBias);
    const int et = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = unit; item < p.n_items; item += n_units) {
      const ItemCoord c = decode_item(p, item);
      const int q_glob = c.qg * BM + q_local;
      const bool active = q_glob < p.nq;
      float theta = active ? -INFINITY : INFINITY;
      int cnt = 0;
      const float* bias = p.bias[c.db];
      const int n_rows = p.n_rows[c.db];
      for (int tile = c.t0 + crank; tile < c.t1; tile += 2) {
        if (bias != nullptr) {
          sbias[acc * BN + et] = bias[static_cast<long long>(tile) * BN + et];
          sbias[acc * BN + 128 + et] = bias[static_cast<long long>(tile) * BN + 128 + et];
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        mbar_wait(tfull_bar(acc), acc_phase, dead, p.err, 0x400u + acc);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
        const int nvalid = n_rows - tile * BN;  // >= BN for full tiles
#pragma unroll 1
        for (int ch = 0; ch < BN / CHUNK; ++ch) {
          uint32_t v[CHUNK];
          tmem_ld32(taddr + ch * CHUNK, v);
          tmem_ld_wait(v);
          if (bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + acc * BN + ch * CHUNK);
#pragma unroll
            for (int j4 = 0; j4 < CHUNK / 4; ++j4) {
              const float4 b = b4[j4];
              v[4 * j4 + 0] = __float_as_uint(__uint_as_float(v[4 * j4 + 0]) + b.x);
              v[4 * j4 + 1] = __float_as_uint(__uint_as_float(v[4 * j4 + 1]) + b.y);
              v[4 * j4 + 2] = __float_as_uint(__uint_as_float(v[4 * j4 + 2]) + b.z);
              v[4 * j4 + 3] = __float_as_uint(__uint_as_float(v[4 * j4 + 3]) + b.w);
            }
          } else if (nvalid < BN) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
              if (ch * CHUNK + j >= nvalid) v[j] = 0xff800000u;  // -inf: zero-filled rows past the end
          }
          const uint32_t idx0 = static_cast<uint32_t>(tile * BN + ch * CHUNK);
          if (p.dump != nullptr && active) {
            float* drow = p.dump + (static_cast<long long>(c.db) * p.nq + q_glob) * p.ld_dump;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
              if (static_cast<int>(idx0) + j < n_rows) drow[idx0 + j] = __uint_as_float(v[j]);
          }
          uint32_t wptr = slot0 + static_cast<uint32_t>(cnt) * 256u;
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) {
            if (__uint_as_float(v[j]) > theta) {
              sts64(wptr, v[j], idx0 + j);
              wptr += 256u;
            }
          }
          cnt = static_cast<int>((wptr - slot0) >> 8);
          if (__any_sync(0xffffffffu, cnt > CAP - CHUNK)) {
            const CandState st = compact_candidates<CAP>(slot0, cnt, theta);
            cnt = st.cnt;
            theta = st.theta;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (__any_sync(0xffffffffu, cnt >= LKEEP)) {
        const CandState st = compact_candidates<CAP>(slot0, cnt, theta);
        cnt = st.cnt;
        theta = st.theta;
      }
      // sub-slice s' = 2 s + cta of database db: [((db * 2S) + s') * n_qt + qt][query][LKEEP]
      const long long oitem =
          (static_cast<long long>(c.db) * (2 * p.S) + 2 * c.s + crank) * p.n_qt + c.qg;
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
  cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  ktimer_end(p.timing, t_start);
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace keds
