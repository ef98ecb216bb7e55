// evidence_kernel.cuh : per-level evidence reductions, conditional resampling, MAP and the
// model-selection epilogue.
//
// Reference: the tail of every level of DLAGP.log_model_evidences (dla_gp.py:157-218):
//   ll -= log S ; NaN-out samples whose absorbers are closer than min_z_separation ;
//   evidence = max + log(nanmean(exp(ll - max))) - level * log S ;
//   W = exp(ll - max), NaN -> 0 ; base_sample_inds[level] = np.random.choice(S, S, p = W / W.sum())
// np.random.choice is searchsorted(cumsum(p) / cumsum(p)[-1], U, 'right') on the next S
// uniforms of the legacy MT19937 stream (numpy/random/mtrand.pyx); W.sum() is NumPy's
// pairwise summation (numpy/core/src/umath/loops_utils.h.src, PW_BLOCKSIZE 128) and cumsum
// is a plain running sum.  Both are reproduced operation for operation so that, given the
// same W, the drawn indices are bit-identical to NumPy's.
#pragma once
#include <stdint.h>
#include <math.h>

namespace dla {

// ---- NumPy pairwise summation, sequential restatement (one thread) -------------------------
__device__ inline double np_pairwise_sum_block(const double* a, int n) {
  // n <= 128 : 8 accumulators, then the fixed combination tree, then the tail
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
  int i;
  for (i = 8; i < n - (n % 8); i += 8) {
    r0 += a[i + 0]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
    r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
  }
  double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  for (; i < n; ++i) res += a[i];
  return res;
}

__device__ inline double np_pairwise_sum(const double* a, int n) {
  // explicit stack instead of recursion: post-order evaluation of the split tree
  // node = (offset, length, state); depth <= 32
  struct Node { int off, len, state; double left; };
  Node st[40];
  int top = 0;
  st[0] = {0, n, 0, 0.0};
  double ret = 0.0;
  while (top >= 0) {
    Node& nd = st[top];
    if (nd.len <= 128) {
      ret = np_pairwise_sum_block(a + nd.off, nd.len);
      --top;
      continue;
    }
    int n2 = nd.len / 2;
    n2 -= n2 % 8;
    if (nd.state == 0) {
      nd.state = 1;
      st[top + 1] = {nd.off, n2, 0, 0.0};
      ++top;
    } else if (nd.state == 1) {
      nd.left = ret;
      nd.state = 2;
      st[top + 1] = {nd.off + n2, nd.len - n2, 0, 0.0};
      ++top;
    } else {
      ret = nd.left + ret;
      --top;
    }
  }
  return ret;
}

// Separation test of one sample at level `level` (dla_gp.py:164-177): the redshifts of its absorbers
// [z_s, z_{b0[s]}, ..., z_{b(level-1)[s]}], sorted; true if any neighbouring pair is closer than the limit.
// Monotone in the level: inserting one more redshift into a sorted chain can only split a gap into smaller ones, so a
// sample that fails at level L fails at every later level.
__device__ __forceinline__ bool absorbers_too_close(const double* __restrict__ z_samples, const int32_t* __restrict__ base_inds,
                                                    int S, int level, int s, double min_z_separation) {
  double zs[9];
  zs[0] = z_samples[s];
  for (int r = 0; r < level; ++r) zs[r + 1] = z_samples[base_inds[(size_t)r * S + s]];
  // insertion sort (<= 9 values) == np.sort along the absorber axis
  for (int a = 1; a <= level; ++a) {
    const double key = zs[a];
    int b = a - 1;
    while (b >= 0 && zs[b] > key) { zs[b + 1] = zs[b]; --b; }
    zs[b + 1] = key;
  }
  bool close = false;
  for (int a = 0; a < level; ++a) close |= (zs[a + 1] - zs[a]) < min_z_separation;
  return close;
}

// Before the likelihoods of level >= 1: the samples the reference overwrites with NaN after computing them
// (separation mask, dla_gp.py:164-177) are known from the redshifts alone, so they are left out of the launch.
// On the bench workload the mask removes 22 % of the level 1-3 evaluations (up to 97 % of level 3 when the resampled
// absorbers pile up on one strong DLA).  The kept samples are compacted, in ascending order, into launch SLOTS:
//   sel[slot]   = sample id                      rows1[slot] = profile-cache row of the newly drawn absorber
//   pos[sample] = slot (or -1)                   rows0[slot] = row of the sample's running product: its own profile
//                                                              at level 1, its slot in the previous level's launch after
// and the likelihood kernel runs over slots (its row indirection rows0 / rows is all it needs); the raw
// log-likelihoods come back in slot order and scatter_ll_kernel puts them at their sample ids, NaN elsewhere.
// A sample kept at level L was kept at level L - 1 (the mask is monotone), so pos_prev[sample] is always a slot.
struct CompactTask {
  const double* z_samples;   // S
  const int32_t* base_inds;  // [(max_dlas-1)][S]
  const int* alive;          // spectrum left the level loop: nothing to evaluate
  const int32_t* pos_prev;   // S : slots of the previous level (level >= 2), nullptr at level 1
  int32_t* sel;              // S out
  int32_t* pos;              // S out
  int32_t* rows0;            // S out
  int32_t* rows1;            // S out
  double* raw_ll;            // S : NaN written for masked samples (the scatter fills the others)
  int* num_sel;              // 1 out
  int S, level;
  double min_z_separation;
};
// grid = num_spectra, block = 1024
__global__ void __launch_bounds__(1024) compact_level_kernel(const CompactTask* __restrict__ tasks) {
  const CompactTask t = tasks[blockIdx.x];
  __shared__ int s_warp[32];  // exclusive offset of every warp within the current chunk of 1024 samples
  __shared__ int s_chunk_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (t.alive && *t.alive == 0) {
    if (tid == 0) *t.num_sel = 0;
    return;
  }
  int base = 0;  // samples kept in the chunks before this one (the same value in every thread)
  for (int s0 = 0; s0 < t.S; s0 += blockDim.x) {
    const int s = s0 + tid;
    bool keep = false;
    if (s < t.S) {
      keep = !absorbers_too_close(t.z_samples, t.base_inds, t.S, t.level, s, t.min_z_separation);
      if (!keep) { t.raw_ll[s] = NAN; t.pos[s] = -1; }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncthreads();  // the previous chunk's offsets have been read
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int inc = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += u;
      }
      s_warp[lane] = inc - v;
      if (lane == 31) s_chunk_total = inc;
    }
    __syncthreads();
    if (keep) {
      const int slot = base + s_warp[warp] + __popc(bal & ((1u << lane) - 1u));  // ascending sample ids
      t.sel[slot] = s;
      t.pos[s] = slot;
      t.rows0[slot] = t.pos_prev ? t.pos_prev[s] : s;
      t.rows1[slot] = t.base_inds[(size_t)(t.level - 1) * t.S + s];
    }
    base += s_chunk_total;
  }
  if (tid == 0) *t.num_sel = base;
}

// raw log-likelihoods of a compacted launch back to their sample ids; grid = (ceil(S / 256), num_spectra)
struct ScatterTask {
  const double* raw_slots;  // [num_sel] in slot order
  const int32_t* sel;       // [num_sel]
  const int* num_sel;       // counts per level; entry `level` is read
  double* raw_ll;           // [S]
};
__global__ void scatter_ll_kernel(const ScatterTask* __restrict__ tasks, int level) {
  const ScatterTask t = tasks[blockIdx.y];
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < t.num_sel[level]) t.raw_ll[t.sel[slot]] = t.raw_slots[slot];
}

// One spectrum-level as the evidence kernel sees it.
struct EvidenceLevel {
  const double* raw_ll;      // S raw log-likelihoods of this level (from the likelihood kernel)
  double* sample_ll;         // destination column: element s at sample_ll[s * ll_stride]
  int ll_stride;             // max_dlas (column of an (S, max_dlas) array) or 1
  const double* z_samples;   // S absorber redshifts (for the separation test)
  const int32_t* base_inds;  // [(max_dlas-1)][S] indices drawn so far (rows < level are valid)
  int32_t* base_out;         // row `level` of base_inds to fill, or nullptr when not resampling
  const double* uniforms;    // S uniforms for this resampling step (or nullptr)
  double* log_evidence;      // one double out
  double* cdf_scratch;       // S doubles of global scratch (W, then p, then cdf)
  int* alive;                // in/out: 0 = this spectrum left the level loop (dla_gp.py:200-206); skip
  int* status;               // set to 2 when a non-final level's evidence is NaN (early exit)
  int S;
  int level;                 // number of additional absorbers (0-based, "num_dlas" of the loop)
  double min_z_separation;
};

// grid = num_spectra, block = 1024
__global__ void __launch_bounds__(1024)
evidence_level_kernel(const EvidenceLevel* __restrict__ levels) {
  const EvidenceLevel lv = levels[blockIdx.x];
  if (lv.alive && *lv.alive == 0) return;
  __shared__ double s_red[32];
  __shared__ double s_red2[32];
  __shared__ int s_cnt[32];
  __shared__ double s_max, s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = lv.S;
  const double logS = log((double)S);

  // ---- ll - log S, separation mask (dla_gp.py:155-177) ------------------------------------
  double tmax = -INFINITY;
  bool any_valid = false;
  for (int s = tid; s < S; s += blockDim.x) {
    double ll = lv.raw_ll[s] - logS;
    if (lv.level > 0 && absorbers_too_close(lv.z_samples, lv.base_inds, S, lv.level, s, lv.min_z_separation)) ll = NAN;
    lv.sample_ll[(size_t)s * lv.ll_stride] = ll;
    if (!isnan(ll)) { tmax = fmax(tmax, ll); any_valid = true; }
  }
  // ---- nanmax ---------------------------------------------------------------------------
  for (int off = 16; off > 0; off >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
  int anyv = __syncthreads_or(any_valid ? 1 : 0);
  if (lane == 0) s_red[warp] = tmax;
  __syncthreads();
  if (tid == 0) {
    double m = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, s_red[w]);
    s_max = anyv ? m : NAN;
  }
  __syncthreads();
  const double maxll = s_max;

  // ---- nanmean(exp(ll - max)) and W ------------------------------------------------------
  double tsum = 0.0;
  int tcnt = 0;
  for (int s = tid; s < S; s += blockDim.x) {
    const double ll = lv.sample_ll[(size_t)s * lv.ll_stride];
    const double pr = exp(ll - maxll);  // NaN stays NaN
    double w = pr;
    if (isnan(pr)) w = 0.0; else { tsum += pr; ++tcnt; }
    if (lv.base_out) lv.cdf_scratch[s] = w;
  }
  for (int off = 16; off > 0; off >>= 1) {
    tsum += __shfl_xor_sync(0xffffffffu, tsum, off);
    tcnt += __shfl_xor_sync(0xffffffffu, tcnt, off);
  }
  if (lane == 0) { s_red2[warp] = tsum; s_cnt[warp] = tcnt; }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    int cnt = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tot += s_red2[w]; cnt += s_cnt[w]; }
    const double mean = cnt > 0 ? tot / (double)cnt : NAN;
    const double ev = maxll + log(mean) - logS * (double)lv.level;  // dla_gp.py:186-190
    *lv.log_evidence = ev;
    if (isnan(ev) && lv.base_out) {  // not the last level: the reference breaks out of the loop here
      if (lv.alive) *lv.alive = 0;
      if (lv.status) *lv.status = 2;
    }
    s_total = ev;
  }
  __syncthreads();
  if (!lv.base_out) return;
  if (isnan(s_total)) return;  // early exit: later levels stay NaN, indices stay 0

  // ---- np.random.choice(S, S, p = W / W.sum()) --------------------------------------------
  double* W = lv.cdf_scratch;
  __threadfence_block();
  __syncthreads();
  if (tid == 0) s_total = np_pairwise_sum(W, S);
  __syncthreads();
  const double wsum = s_total;
  for (int s = tid; s < S; s += blockDim.x) W[s] = W[s] / wsum;  // p
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;  // p.cumsum(): running sum in index order
    for (int s = 0; s < S; ++s) { run += W[s]; W[s] = run; }
    s_total = run;
  }
  __syncthreads();
  const double last = s_total;
  for (int s = tid; s < S; s += blockDim.x) W[s] = W[s] / last;  // cdf /= cdf[-1]
  __syncthreads();
  for (int s = tid; s < S; s += blockDim.x) {
    const double u = lv.uniforms[s];
    // searchsorted(cdf, u, side='right'): number of entries <= u
    int lo = 0, hi = S;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (W[mid] <= u) lo = mid + 1; else hi = mid;
    }
    lv.base_out[s] = lo;
  }
}

// ---- resampling alone (for dla_resample_indices and its bit-exactness tests) ---------------
__global__ void __launch_bounds__(1024)
resample_kernel(const double* __restrict__ Win, const double* __restrict__ uniforms, int S, double* scratch,
                int32_t* out) {
  __shared__ double s_total;
  const int tid = threadIdx.x;
  for (int s = tid; s < S; s += blockDim.x) scratch[s] = Win[s];
  __syncthreads();
  if (tid == 0) s_total = np_pairwise_sum(scratch, S);
  __syncthreads();
  const double wsum = s_total;
  for (int s = tid; s < S; s += blockDim.x) scratch[s] = scratch[s] / wsum;
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int s = 0; s < S; ++s) { run += scratch[s]; scratch[s] = run; }
    s_total = run;
  }
  __syncthreads();
  const double last = s_total;
  for (int s = tid; s < S; s += blockDim.x) scratch[s] = scratch[s] / last;
  __syncthreads();
  for (int s = tid; s < S; s += blockDim.x) {
    const double u = uniforms[s];
    int lo = 0, hi = S;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (scratch[mid] <= u) lo = mid + 1; else hi = mid;
    }
    out[s] = lo;
  }
}

// ---- a13: DLAGP.maximum_a_posteriori (dla_gp.py:428-472) ------------------------------------
struct MapTask {
  const double* sample_ll;   // (S, max_dlas) row-major
  const int32_t* base_inds;  // (max_dlas-1, S)
  const double* z_samples;   // S
  const double* log_nhi;     // S
  double* map_z;             // (max_dlas, max_dlas) NaN padded
  double* map_log_nhi;       // (max_dlas, max_dlas)
  int32_t* map_ind;          // max_dlas argmax indices (-1: all-NaN column, where np.nanargmax raises)
  int S, max_dlas;
};

// grid = (max_dlas, num_spectra), block = 256: first maximum wins (np.nanargmax)
__global__ void __launch_bounds__(256)
map_kernel(const MapTask* __restrict__ tasks) {
  const MapTask t = tasks[blockIdx.y];
  const int level = blockIdx.x;
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  double best = -INFINITY;
  int bidx = -1;
  for (int s = threadIdx.x; s < t.S; s += blockDim.x) {
    const double v = t.sample_ll[(size_t)s * t.max_dlas + level];
    if (!isnan(v) && (bidx < 0 || v > best)) { best = v; bidx = s; }
  }
  s_val[threadIdx.x] = best;
  s_idx[threadIdx.x] = bidx;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const double ov = s_val[threadIdx.x + off];
      const int oi = s_idx[threadIdx.x + off];
      const int mi = s_idx[threadIdx.x];
      if (oi >= 0 && (mi < 0 || ov > s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < mi))) {
        s_val[threadIdx.x] = ov;
        s_idx[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int mi = s_idx[0];
    if (t.map_ind) t.map_ind[level] = mi;
    for (int c = 0; c < t.max_dlas; ++c) {
      double z = NAN, ln = NAN;
      if (mi >= 0 && c <= level) {
        const int idx = c == 0 ? mi : t.base_inds[(size_t)(c - 1) * t.S + mi];
        z = t.z_samples[idx];
        ln = t.log_nhi[idx];
      }
      t.map_z[level * t.max_dlas + c] = z;
      t.map_log_nhi[level * t.max_dlas + c] = ln;
    }
  }
}

// ---- a12: BayesModelSelect.model_selection + posteriors (bayesian_model_selection.py:48-149)
// one thread per spectrum; m = 2 + max_dlas models [null, subDLA, DLA 1..max]
__device__ inline double logsumexp_dev(const double* a, int n) {
  double mx = -INFINITY;
  for (int i = 0; i < n; ++i) if (a[i] > mx) mx = a[i];   // scipy: max over finite values
  if (!isfinite(mx)) mx = 0.0;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += exp(a[i] - mx);
  return log(s) + mx;
}

__global__ void model_selection_kernel(const double* __restrict__ log_priors_in, const double* __restrict__ log_lik,
                                       int num_spectra, int max_dlas, double* log_priors, double* log_post,
                                       double* model_post, double* p_dla, double* p_no_dla) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= num_spectra) return;
  const int m = 2 + max_dlas;
  double pri[16], post[16];
  for (int i = 0; i < m; ++i) pri[i] = log_priors_in[(size_t)q * m + i];
  pri[0] = log(1.0 - exp(logsumexp_dev(pri + 1, m - 1)));  // :79-80
  for (int i = 0; i < m; ++i) post[i] = log_lik[(size_t)q * m + i] + pri[i];
  const double lse = logsumexp_dev(post, m);
  double pd = 0.0;
  for (int i = 0; i < m; ++i) {
    const double mp = exp(post[i] - lse);
    if (log_priors) log_priors[(size_t)q * m + i] = pri[i];
    if (log_post) log_post[(size_t)q * m + i] = post[i];
    if (model_post) model_post[(size_t)q * m + i] = mp;
    if (i >= 2) pd += mp;  // :141-145 (np.sum over the last max_dlas entries)
  }
  if (p_dla) p_dla[q] = pd;
  if (p_no_dla) p_no_dla[q] = 1.0 - pd;
}

}  // namespace dla
