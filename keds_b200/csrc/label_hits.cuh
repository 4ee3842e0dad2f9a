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
template <int RIF, int MINB>
__global__ void __launch_bounds__(HITS_THREADS, MINB)
k_select_hits(const HitsParams p) {
  extern __shared__ uint8_t hs_smem[];
  const int q = blockIdx.x;
  const int qt = q / BM, ql = q % BM;
  const int slots = p.S * LKEEP;
  float* qvec = reinterpret_cast<float*>(hs_smem);                          // d (16-B aligned)
  unsigned int* keys = reinterpret_cast<unsigned int*>(qvec + ((p.d + 3) & ~3));  // slots
  unsigned int* ids = keys + slots;                                         // slots
  unsigned int* c_key = ids + slots;                                        // rmax  the set C = { a >= a_(kmax) - 2 eps }
  unsigned int* c_id = c_key + p.rmax;                                      // rmax
  float* c_sc = reinterpret_cast<float*>(c_id + p.rmax);                    // rmax  exact score (band rows only)
  unsigned int* c_info = reinterpret_cast<unsigned int*>(c_sc + p.rmax);    // rmax  bit j: in cut j's band; bit 8: label match; bit 16 + j: certainly in cut j
  unsigned int* amb = c_info + p.rmax;                                      // rmax  indices of the rows to re-score
  unsigned int* hist = amb + p.rmax;                                        // 256
  float* red = reinterpret_cast<float*>(hist + 256);                        // 32
  unsigned int* bcast = reinterpret_cast<unsigned int*>(red + 32);          // 4
  int* counters = reinterpret_cast<int*>(bcast + 4);                        // 4
  unsigned int* kth = reinterpret_cast<unsigned int*>(counters + 4);        // HITS_MAX_CUTS
  int* n_cert = reinterpret_cast<int*>(kth + HITS_MAX_CUTS);                // HITS_MAX_CUTS
  int* n_hit = n_cert + HITS_MAX_CUTS;                                      // HITS_MAX_CUTS

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (tid < 4) counters[tid] = 0;
  if (tid < HITS_MAX_CUTS) {
    n_cert[tid] = 0;
    n_hit[tid] = 0;
  }
  for (int c = tid; c < p.d; c += blockDim.x) qvec[c] = p.q_f32[static_cast<long long>(q) * p.d + c];
  const float xb = __uint_as_float(p.dbstat[0]);
  const float xd = __uint_as_float(p.dbstat[1]);
  const float xn2 = __uint_as_float(p.dbstat[2]);
  const long long my_label = p.qlabel[q];
  griddep_wait();  // candidate lists (and qstat from k_prep_rows) are visible from here on
  const unsigned long long t_start = ktimer_begin(p.timing);

  // ---- A: every list's candidate line, count and threshold (half a warp per list)
  const float4 qs = p.qstat[q];
  float th_max = -INFINITY;
  int n_valid = 0;
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
        keys[s * LKEEP + e] = id != PAD_ID ? f32_to_key(sc) : 0u;
        ids[s * LKEEP + e] = id;
      }
      th_max = fmaxf(th_max, th[it]);
      n_valid += cn[it];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    th_max = fmaxf(th_max, __shfl_xor_sync(0xffffffffu, th_max, o));
    n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
  }
  if (lane == 0) {
    red[warp] = th_max;
    atomicAdd(&counters[3], n_valid);
  }
  __syncthreads();
  th_max = red[0];
  for (int w = 1; w < nwarps; ++w) th_max = fmaxf(th_max, red[w]);
  const int n = counters[3];

  // the search's error bound (rerank.cuh)
  const float qn = sqrtf(qs.x);
  const int d_pad = (p.d + BK - 1) / BK * BK;
  float eps = qs.z * xb + qn * xd + (static_cast<float>(d_pad) * 2.4e-7f) * qs.y * xb;
  if (p.metric == METRIC_L2) eps += 1.0e-6f * xn2 + 1.2e-7f * (qs.y * xb + 0.5f * xn2);
  eps *= 1.0001f * p.eps_scale;
  if (!(eps == eps)) eps = INFINITY;
  const float band = 2.f * eps;

  // ---- B: a_(kmax), the certificate, the set C
  const int kmax = p.ks[p.nks - 1];
  bool ok = n >= kmax;
  int m = 0;
  if (ok) {
    const unsigned int kk = block_kth_largest(keys, slots, kmax, hist, bcast);  // padding keys are 0: never among the top n
    const float tau = key_to_f32(kk) - band;
    for (int i0 = 0; i0 < slots; i0 += blockDim.x) {
      const int i = i0 + tid;
      unsigned int id = PAD_ID, key = 0u;
      if (i < slots) {
        id = ids[i];
        key = keys[i];
      }
      const bool hit = id != PAD_ID && key_to_f32(key) >= tau;
      const unsigned int bal = __ballot_sync(0xffffffffu, hit);
      if (bal != 0u) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < p.rmax) {
          c_key[pos] = key;
          c_id[pos] = id;
        }
      }
    }
    if (tid == 0) kth[p.nks - 1] = kk;
    __syncthreads();
    m = counters[0];
    ok = (m <= p.rmax) && (th_max == -INFINITY || th_max < tau);
  }
  if (!ok) {
    if (tid == 0) {
      const int pos = atomicAdd(p.n_flagged, 1);
      p.flagged[pos] = q;
      if (p.band_max != nullptr && m > 2 * kmax) atomicMax(p.band_max, static_cast<unsigned int>(m));
    }
    ktimer_end(p.timing, t_start);
    return;
  }
  if (tid == 0 && p.band_max != nullptr && m > 2 * kmax) atomicMax(p.band_max, static_cast<unsigned int>(m));
  // the other cut points: c-th largest approximate score, all of them inside C (ks[j] <= kmax <= m);
  // one rank-counting pass over C serves every cut (equal keys write the same value)
  for (int i = tid; i < m; i += blockDim.x) {
    const unsigned int mine = c_key[i];
    int gt = 0, ge = 0;
#pragma unroll 8
    for (int o = 0; o < m; ++o) {
      const unsigned int other = c_key[o];
      gt += other > mine;
      ge += other >= mine;
    }
    for (int j = 0; j < p.nks - 1; ++j)
      if (gt < p.ks[j] && ge >= p.ks[j]) kth[j] = mine;
  }
  __syncthreads();

  // ---- C: classify every row of C against every cut; labels; the rows that need an exact score
  for (int i0 = 0; i0 < m; i0 += blockDim.x) {
    const int i = i0 + tid;
    const bool live = i < m;
    const float a = live ? key_to_f32(c_key[i]) : -INFINITY;
    unsigned int info = (live && p.row_labels[c_id[i]] == my_label) ? 0x100u : 0u;
    for (int j = 0; j < p.nks; ++j) {
      const float ac = key_to_f32(kth[j]);
      const bool cert = live && a > ac + band;
      if (cert) info |= 1u << (16 + j);
      else if (live && a >= ac - band) info |= 1u << j;
      // one shared-memory atomic per warp and cut
      const unsigned int bc = __ballot_sync(0xffffffffu, cert);
      const unsigned int bh = __ballot_sync(0xffffffffu, cert && (info & 0x100u));
      if (lane == 0 && bc != 0u) {
        atomicAdd(&n_cert[j], __popc(bc));
        if (bh != 0u) atomicAdd(&n_hit[j], __popc(bh));
      }
    }
    if (live) c_info[i] = info;
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

  // ---- D: exact fp32 scores of the band rows, three rows per warp in flight
  for (int c0 = warp * RIF; c0 < n_amb; c0 += nwarps * RIF) {
    int ii[RIF];
    const float* xr[RIF];
#pragma unroll
    for (int r = 0; r < RIF; ++r) {
      ii[r] = static_cast<int>(amb[min(c0 + r, n_amb - 1)]);
      xr[r] = p.x_f32 + static_cast<long long>(c_id[ii[r]]) * p.d;
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

  // ---- E: per cut, the band rows compete for the places the certain rows leave
  for (int j = 0; j < p.nks; ++j) {
    const int places = p.ks[j] - n_cert[j];
    const unsigned int bit = 1u << j;
    for (int c = tid; c < n_amb; c += blockDim.x) {
      const int i = static_cast<int>(amb[c]);
      const unsigned int info = c_info[i];
      if (!(info & bit) || !(info & 0x100u)) continue;   // only a label match can add a hit
      const unsigned long long mine = order_key(c_sc[i], c_id[i]);
      int beats = 0;
      for (int o = 0; o < n_amb; ++o) {
        const int io = static_cast<int>(amb[o]);
        beats += (c_info[io] & bit) && order_key(c_sc[io], c_id[io]) > mine;
      }
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
