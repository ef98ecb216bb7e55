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
    if (lv.level > 0) {
      // z of all absorbers of this sample: [z_s, z_{b0[s]}, ...]; NaN if any pair closer than the limit
      double zs[9];
      zs[0] = lv.z_samples[s];
      for (int r = 0; r < lv.level; ++r) zs[r + 1] = lv.z_samples[lv.base_inds[(size_t)r * S + s]];
      // insertion sort (<= 9 values) == np.sort along the absorber axis
      for (int a = 1; a <= lv.level; ++a) {
        const double key = zs[a];
        int b = a - 1;
        while (b >= 0 && zs[b] > key) { zs[b + 1] = zs[b]; --b; }
        zs[b + 1] = key;
      }
      bool close = false;
      for (int a = 0; a < lv.level; ++a) close |= (zs[a + 1] - zs[a]) < lv.min_z_separation;
      if (close) ll = NAN;
    }
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
