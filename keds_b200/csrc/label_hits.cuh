// Label hits at several cut points straight from the candidate lists: the counting core of
// get_metrics_imgnet (src/eval_utils.py:1101-1118) without the exact re-score of all k rows.
//
// The reference needs, per query, hits_c = #{ rows among the exact top-c with the query's label }
// for c in {1, 5, 10, 50, 100, 200}; neither distances nor the order inside a top-c set matter.
// A top-200 SEARCH has to read 200+ fp32 rows per query to hand back exact distances (6 GB per
// 10,000 queries, twice the GEMM's time at the 50k-gallery shape). Membership is cheaper. With
// a(n) the 16-bit-operand score, s(n) the exact one, |a - s| <= eps, and a_(c) the c-th largest a:
//   a(n) >  a_(c) + 2 eps   =>  n is in the exact top-c  (fewer than c rows can score that high)
//   a(n) <  a_(c) - 2 eps   =>  n is not                 (c rows score higher for certain)
// so only the rows inside the +-2 eps band around a cut point need exact fp32 scores: the band rows
// of cut c compete, by exact score (then lower row id), for the c - #certain places that are left.
// About 65 of the ~250 candidate rows at the 50k shape. The certificate of the search (every
// list's drop threshold below a_(kmax) - 2 eps) is checked the same way; a query that fails it is
// queued for the exact fallback, whose result row is then counted by k_hits_from_rows.
#pragma once

namespace keds {

constexpr int HITS_MAX_CUTS = 8;
constexpr int HITS_THREADS = 128;

struct HitsParams {
  int n_qt, S, nq, d, metric, nks, rmax;
  int ks[HITS_MAX_CUTS];       // ascending cut points, ks[nks - 1] = kmax
  const uint2* cand;           // [list][BM][LKEEP] (score_topk_sm100.cuh)
  const int* cand_cnt;
  const float* cand_theta;
  const float* q_f32;          // [nq][d]
  const float4* qstat;         // [nq] {|q|^2, |q16|, |q - q16|}
  const float* x_f32;
  const unsigned int* dbstat;
  const long long* row_labels; // [ntotal]
  const long long* qlabel;     // [nq]
  int* hits;                   // [nq][nks]
  int* flagged;
  int* n_flagged;
  float eps_scale;
  unsigned int* band_max;
  unsigned long long* timing;
};

// RIF = band rows per warp in flight during the exact re-score, MINB = blocks per SM the register
// budget is cut for.
// Locate the histogram bin that holds the need-th largest key: the largest b with
// count(bins >= b) >= need. 1024 bins, warp 0 works (lane l owns bins 32 l .. 32 l + 31); uniform
// call, ends with a __syncthreads. bcast[0] = b, bcast[3] = count(bins > b).
__device__ __forceinline__ void block_find_bin1024(const unsigned int* hist, int need, unsigned int* bcast) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned int tot = 0;
    for (int i = 0; i < 32; ++i) tot += hist[lane * 32 + i];
    unsigned int suf = tot;  // inclusive suffix sum over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += up;
    }
    unsigned int above = suf - tot;
    if (static_cast<int>(above) < need && static_cast<int>(suf) >= need) {
      for (int i = 31; i >= 0; --i) {
        const unsigned int h = hist[lane * 32 + i];
        if (static_cast<int>(above + h) >= need) {
          bcast[0] = static_cast<unsigned int>(lane * 32 + i);
          bcast[3] = above;
          break;
        }
        above += h;
      }
    }
  }
  __syncthreads();
}

// RIF = band rows per warp in flight during the exact re-score, MINB = blocks per SM the register
// budget is cut for.
template <int RIF, int MINB>
__global__ void __launch_bounds__(HITS_THREADS, MINB)
k_select_hits(const HitsParams p) {
  extern __shared__ uint8_t hs_smem[];
  const int q = blockIdx.x;
  const int qt = q / BM, ql = q % BM;
  const int slots = p.S * LKEEP;
  float* qvec = reinterpret_cast<float*>(hs_smem);                                  // d (16-B aligned)
  unsigned long long* cs = reinterpret_cast<unsigned long long*>(qvec + ((p.d + 3) & ~3));  // rmax  the set C, sorted: key << 32 | row id
  unsigned int* keys = reinterpret_cast<unsigned int*>(cs + p.rmax);                // slots
  unsigned int* ids = keys + slots;                                                 // slots
  float* c_sc = reinterpret_cast<float*>(ids + slots);                              // rmax  exact score (band rows only)
  unsigned int* c_info = reinterpret_cast<unsigned int*>(c_sc + p.rmax);            // rmax  bit j: in cut j's band; bit 8: label match
  unsigned int* amb = c_info + p.rmax;                                              // rmax  sorted positions of the rows to re-score
  unsigned int* hist = amb + p.rmax;                                                // 1024
  float* red = reinterpret_cast<float*>(hist + 1024);                               // 32
  unsigned int* bcast = reinterpret_cast<unsigned int*>(red + 32);                  // 4
  int* counters = reinterpret_cast<int*>(bcast + 4);                                // 4
  int* n_cert = counters + 4;                                                       // HITS_MAX_CUTS  rows certainly inside cut j (= start of its band in sorted order)
  int* band_end = n_cert + HITS_MAX_CUTS;                                           // HITS_MAX_CUTS  end of cut j's band
  int* n_hit = band_end + HITS_MAX_CUTS;                                            // HITS_MAX_CUTS

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (tid < 4) counters[tid] = 0;
  if (tid < HITS_MAX_CUTS) n_hit[tid] = 0;
  if (tid == 0) {
    bcast[0] = 0u;           // histogram bin of the cut (0: everything is a candidate)
    bcast[1] = 0xFFFFFFFFu;  // smallest valid key
    bcast[2] = 0u;           // largest
    bcast[3] = 0u;           // candidates above that bin
  }
  for (int i = tid; i < 1024; i += blockDim.x) hist[i] = 0u;
  for (int c = tid; c < p.d; c += blockDim.x) qvec[c] = p.q_f32[static_cast<long long>(q) * p.d + c];
  const float xb = __uint_as_float(p.dbstat[0]);
  const float xd = __uint_as_float(p.dbstat[1]);
  const float xn2 = __uint_as_float(p.dbstat[2]);
  const long long my_label = p.qlabel[q];
  __syncthreads();
  griddep_wait();  // candidate lists (and qstat from k_prep_rows) are visible from here on
  const unsigned long long t_start = ktimer_begin(p.timing);

  // ---- A: every list's candidate line, count and threshold (half a warp per list); key range
  const float4 qs = p.qstat[q];
  float th_max = -INFINITY;
  int n_valid = 0;
  unsigned int kmin = 0xFFFFFFFFu, kmaxk = 0u;
  constexpr int MAX_IT = 6;
  const int e = lane & 15;
  for (int sb = 0; sb < p.S; sb += MAX_IT * nwarps * 2) {
    uint2 en[MAX_IT];
    float th[MAX_IT];
    int cn[MAX_IT];
#pragma unroll
    for (int it = 0; it < MAX_IT; ++it) {
      const int s = sb + (it * nwarps + warp) * 2 + (lane >> 4);
      en[it] = make_uint2(0xff800000u, PAD_ID);
      th[it] = -INFINITY;
      cn[it] = 0;
      if (s < p.S) {
        const long long item = static_cast<long long>(s) * p.n_qt + qt;
        en[it] = __ldcg(p.cand + (item * BM + ql) * LKEEP + e);
        if (e == 0) {
          th[it] = __ldcg(p.cand_theta + item * BM + ql);
          cn[it] = __ldcg(p.cand_cnt + item * BM + ql);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < MAX_IT; ++it) {
      const int s = sb + (it * nwarps + warp) * 2 + (lane >> 4);
      float sc = __uint_as_float(en[it].x);
      if (sc == 0.f) sc = 0.f;
      const unsigned int id = en[it].y;
      if (s < p.S) {
        const unsigned int key = id != PAD_ID ? f32_to_key(sc) : 0u;
        keys[s * LKEEP + e] = key;
        ids[s * LKEEP + e] = id;
        if (id != PAD_ID) {
          kmin = min(kmin, key);
          kmaxk = max(kmaxk, key);
        }
      }
      th_max = fmaxf(th_max, th[it]);
      n_valid += cn[it];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    th_max = fmaxf(th_max, __shfl_xor_sync(0xffffffffu, th_max, o));
    n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmaxk = max(kmaxk, __shfl_xor_sync(0xffffffffu, kmaxk, o));
  }
  if (lane == 0) {
    red[warp] = th_max;
    atomicAdd(&counters[3], n_valid);
    atomicMin(&bcast[1], kmin);
    atomicMax(&bcast[2], kmaxk);
  }
  __syncthreads();
  th_max = red[0];
  for (int w = 1; w < nwarps; ++w) th_max = fmaxf(th_max, red[w]);
  const int n = counters[3];
  kmin = bcast[1];
  kmaxk = bcast[2];

  // the search's error bound (rerank.cuh)
  const float qn = sqrtf(qs.x);
  const int d_pad = (p.d + BK - 1) / BK * BK;
  float eps = qs.z * xb + qn * xd + (static_cast<float>(d_pad) * 2.4e-7f) * qs.y * xb;
  if (p.metric == METRIC_L2) eps += 1.0e-6f * xn2 + 1.2e-7f * (qs.y * xb + 0.5f * xn2);
  eps *= 1.0001f * p.eps_scale;
  if (!(eps == eps)) eps = INFINITY;
  const float band = 2.f * eps;

  // ---- B: a lower bound of a_(kmax) from ONE histogram pass over a linear map of the key range
  // (one query's candidate scores span a narrow range: 1024 bins resolve it ~10x finer than the
  // band), then the set C' = { a >= bound - 2 eps }, a superset of C = { a >= a_(kmax) - 2 eps }
  const int kcut = p.ks[p.nks - 1];
  bool ok = n >= kcut;
  int m = 0;
  if (ok) {
    const float lo1 = key_to_f32(kmin), hi1 = key_to_f32(kmaxk);
    const float inv1 = hi1 > lo1 ? 1023.0f / (hi1 - lo1) : 0.f;
    for (int i = tid; i < slots; i += blockDim.x) {
      if (ids[i] != PAD_ID) {
        const unsigned int b = min(1023u, static_cast<unsigned int>((key_to_f32(keys[i]) - lo1) * inv1));
        atomicAdd(&hist[b], 1u);
      }
    }
    __syncthreads();
    block_find_bin1024(hist, kcut, bcast);
    // every score in a bin >= b* is at least the nominal lower edge of bin b* - 1 (a whole bin of
    // slack for the float rounding of the map); at least kcut scores are, so a_(kmax) is too
    const unsigned int bstar = bcast[0];
    const int above1 = static_cast<int>(bcast[3]);
    float bound = (bstar >= 1u && inv1 > 0.f) ? lo1 + static_cast<float>(bstar - 1u) / inv1 : lo1;
    if (inv1 > 0.f && 1.0f / inv1 > 0.5f * band) {
      // An outlier (a near-duplicate of the query; under L2 a score on the other side of zero)
      // stretches the range until one bin is as wide as the band: resolve bin b* with a second
      // histogram over its own range. Membership is decided by the same map as in the first pass,
      // so the count above the bin carries over exactly.
      __syncthreads();
      for (int i = tid; i < 1024; i += blockDim.x) hist[i] = 0u;
      __syncthreads();
      const float lo2 = lo1 + (static_cast<float>(bstar) - 0.01f) / inv1;
      const float inv2 = 1023.0f * inv1 / 1.02f;
      for (int i = tid; i < slots; i += blockDim.x) {
        if (ids[i] != PAD_ID) {
          const float sc = key_to_f32(keys[i]);
          if (min(1023u, static_cast<unsigned int>((sc - lo1) * inv1)) == bstar) {
            const float f = fminf(fmaxf((sc - lo2) * inv2, 0.f), 1023.f);
            atomicAdd(&hist[static_cast<unsigned int>(f)], 1u);
          }
        }
      }
      __syncthreads();
      block_find_bin1024(hist, kcut - above1, bcast);
      const unsigned int b2 = bcast[0];
      if (b2 >= 1u) bound = fmaxf(bound, lo2 + static_cast<float>(b2 - 1u) / inv2);
    }
    const float tau0 = bound - band;
    for (int i0 = 0; i0 < slots; i0 += blockDim.x) {
      const int i = i0 + tid;
      unsigned int id = PAD_ID, key = 0u;
      if (i < slots) {
        id = ids[i];
        key = keys[i];
      }
      const bool hit = id != PAD_ID && key_to_f32(key) >= tau0;
      const unsigned int bal = __ballot_sync(0xffffffffu, hit);
      if (bal != 0u) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < p.rmax) cs[pos] = (static_cast<unsigned long long>(key) << 32) | id;
      }
    }
    __syncthreads();
    m = counters[0];
    ok = m <= p.rmax && m >= kcut;
  }
  float tau = -INFINITY;
  if (ok) {
    // ---- C: sort C' by approximate score, best first (bitonic, padded to a power of two with zeros)
    int P = 64;
    while (P < m) P <<= 1;
    for (int i = m + tid; i < P; i += blockDim.x) cs[i] = 0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (P >> 1); t += blockDim.x) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int l = i | j;
          const unsigned long long a = cs[i], b = cs[l];
          if ((a < b) == ((i & k) == 0)) {
            cs[i] = b;
            cs[l] = a;
          }
        }
        __syncthreads();
      }
    }
    // the certificate: no list dropped a row scoring a_(kmax) - 2 eps or more
    tau = key_to_f32(static_cast<unsigned int>(cs[kcut - 1] >> 32)) - band;
    ok = th_max == -INFINITY || th_max < tau;
  }
  if (!ok) {
    if (tid == 0) {
      const int pos = atomicAdd(p.n_flagged, 1);
      p.flagged[pos] = q;
      if (p.band_max != nullptr && m > 2 * kcut) atomicMax(p.band_max, static_cast<unsigned int>(m));
    }
    ktimer_end(p.timing, t_start);
    return;
  }

  // ---- D: per cut, where its certain rows end and where its band ends in the sorted order
  // (descending scores: two binary searches by one thread per cut)
  if (tid < p.nks) {
    const float ac = key_to_f32(static_cast<unsigned int>(cs[p.ks[tid] - 1] >> 32));
    const float hi = ac + band, lo = ac - band;
    int a0 = 0, a1 = m;  // first position with a <= hi
    while (a0 < a1) {
      const int mid = (a0 + a1) >> 1;
      if (key_to_f32(static_cast<unsigned int>(cs[mid] >> 32)) > hi) a0 = mid + 1; else a1 = mid;
    }
    n_cert[tid] = a0;
    int b0 = a0, b1 = m;  // first position with a < lo
    while (b0 < b1) {
      const int mid = (b0 + b1) >> 1;
      if (key_to_f32(static_cast<unsigned int>(cs[mid] >> 32)) >= lo) b0 = mid + 1; else b1 = mid;
    }
    band_end[tid] = b0;
  }
  __syncthreads();
  const int m_used = band_end[p.nks - 1];  // rows behind the last band are out of every cut
  if (tid == 0 && p.band_max != nullptr && m_used > 2 * kcut) atomicMax(p.band_max, static_cast<unsigned int>(m_used));
  for (int i0 = 0; i0 < m_used; i0 += blockDim.x) {
    const int i = i0 + tid;
    const bool live = i < m_used;
    unsigned int info = 0u;
    if (live) {
      if (p.row_labels[static_cast<unsigned int>(cs[i])] == my_label) info = 0x100u;
      for (int j = 0; j < p.nks; ++j) {
        if (i >= n_cert[j] && i < band_end[j]) info |= 1u << j;
        else if (i < n_cert[j] && info >= 0x100u) atomicAdd(&n_hit[j], 1);  // a certain row with the query's label (rare)
      }
      c_info[i] = info;
    }
    const bool need = live && (info & 0xffu) != 0u;
    const unsigned int bn = __ballot_sync(0xffffffffu, need);
    if (bn != 0u) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&counters[1], __popc(bn));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (need) amb[base + __popc(bn & ((1u << lane) - 1u))] = static_cast<unsigned int>(i);
    }
  }
  __syncthreads();
  const int n_amb = counters[1];

  // ---- E: exact fp32 scores of the band rows, RIF rows per warp in flight
  for (int c0 = warp * RIF; c0 < n_amb; c0 += nwarps * RIF) {
    int ii[RIF];
    const float* xr[RIF];
#pragma unroll
    for (int r = 0; r < RIF; ++r) {
      ii[r] = static_cast<int>(amb[min(c0 + r, n_amb - 1)]);
      xr[r] = p.x_f32 + static_cast<long long>(static_cast<unsigned int>(cs[ii[r]])) * p.d;
    }
    float sc[RIF];
    warp_exact_score_multi<RIF>(qvec, xr, p.d, p.metric, lane, sc);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RIF; ++r)
        if (c0 + r < n_amb) c_sc[ii[r]] = p.metric == METRIC_L2 ? -sc[r] : sc[r];   // larger is better from here on
    }
  }
  __syncthreads();

  // ---- F: per cut, the band rows compete for the places the certain rows leave
  for (int j = 0; j < p.nks; ++j) {
    const int places = p.ks[j] - n_cert[j];
    const int b0 = n_cert[j], b1 = band_end[j];
    for (int i = b0 + tid; i < b1; i += blockDim.x) {
      if (!(c_info[i] & 0x100u)) continue;   // only a label match can add a hit
      const unsigned long long mine = order_key(c_sc[i], static_cast<unsigned int>(cs[i]));
      int beats = 0;
      for (int o = b0; o < b1; ++o) beats += order_key(c_sc[o], static_cast<unsigned int>(cs[o])) > mine;
      if (beats < places) atomicAdd(&n_hit[j], 1);
    }
  }
  __syncthreads();
  if (tid < p.nks) p.hits[static_cast<long long>(q) * p.nks + tid] = n_hit[tid];
  ktimer_end(p.timing, t_start);
}

// hits of the queued queries from their exact result rows (written by k_exact_fallback)
__global__ void k_hits_from_rows(const int* __restrict__ flagged, const int* __restrict__ n_flagged,
                                 const long long* __restrict__ I, int kmax, long long id_offset,
                                 const long long* __restrict__ row_labels,
                                 const long long* __restrict__ qlabel, HitsParams p) {
  griddep_wait();
  const int nf = *n_flagged;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    const int q = flagged[f];
    const long long lab = qlabel[q];
    int acc = 0, ki = 0;
    for (int j = 0; j < kmax && ki < p.nks; ++j) {
      const long long id = I[static_cast<long long>(q) * kmax + j];
      acc += (id >= 0 && row_labels[id - id_offset] == lab);
      while (ki < p.nks && j + 1 == p.ks[ki]) p.hits[static_cast<long long>(q) * p.nks + ki++] = acc;
    }
    while (ki < p.nks) p.hits[static_cast<long long>(q) * p.nks + ki++] = acc;
  }
}

}  // namespace keds
