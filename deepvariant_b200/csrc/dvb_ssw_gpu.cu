// Batched Smith-Waterman on the device (SURVEY.md 8(f) "next" row #3: read-to-haplotype alignment of the alt-aligned pileups and of the
// realigner - deepvariant/realigner/ssw.h over libssw, fast_pass_aligner.cc:SswAlignReadsToHaplotypes, alt_aligned_pileup_lib.cc:278-313).
//
// libssw's result is two full-matrix scans plus a banded traceback (csrc/dvb_ssw.cu restates them with the library's tie-breaking):
//   scan 1  H over (query x reference): score, FIRST reference column holding the global maximum, SMALLEST query row in that column;
//   scan 2  the same recurrences over the reversed query prefix, walking the reference backwards from that column, stopping at the
//           first column whose maximum equals the score: the alignment's begin;
//   banded  traceback inside [begin, end] x [begin, end] -> CIGAR.
// The two scans are the O(|query| x |reference|) part.  Here ONE WARP per alignment runs them as an anti-diagonal wavefront: lane l owns
// query row 32 s + l + 1 of strip s and processes reference column t - l at step t; H(i-1, j-1) and F(i, j) come from the lane above
// by shuffle, E(i, j) stays in the lane's registers, the last row of a strip leaves H / F per column in shared memory for the next
// strip; the reference codes sit in shared memory.  Both of libssw's "first column, smallest row" rules are order-free statements
// (lexicographic minimum of (-H, column in scan order, row)), so the wavefront order does not change the answer - and scan 2's early
// stop is the same statement, because no cell of the reversed problem can exceed the score.  The banded traceback (O(|query| x band))
// then runs on the host inside the window the scans delimit (dvb_ssw_internal::FinishFromEnds, shared with dvb_ssw_align).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "dvb_common.h"

namespace dvb_ssw_internal {
int FinishFromEnds(const int8_t* r, const int8_t* q, int query_len, const int8_t* mat, int gap_open, int gap_extend, DvbSswAlignment* out,
                   char* cigar_out, int64_t cigar_cap);
int8_t BaseCode(char c);
}  // namespace dvb_ssw_internal

namespace {

constexpr int kSswMaxRef = 2048;          // reference columns a warp keeps in shared memory (longer: host path)
constexpr int kSswWarps = 2;              // alignments per CTA

struct SswJob { long long ref_off, q_off; int ref_len, q_len; };
struct SswEnds { int score, ref_end, q_end, ref_begin, q_begin, status; };

struct WarpSmem {
  int8_t ref[kSswMaxRef];
  short hb[2][kSswMaxRef];
  short fb[2][kSswMaxRef];
};

struct Best3 { int h, col, row; };
__device__ __forceinline__ bool better(const Best3& a, const Best3& b) {      // a beats b
  return a.h > b.h || (a.h == b.h && (a.col < b.col || (a.col == b.col && a.row < b.row)));
}

// One scan: columns c = 0 .. ncols-1 are reference positions ref0 + c * rstep (codes in sm.ref), rows i = 1 .. qlen are query positions
// q[q0 + (i - 1) * qstep].  Returns the lexicographically best cell (H desc, column asc, row asc); h == 0 when nothing scores.
__device__ Best3 WarpScan(WarpSmem& sm, int ref0, int rstep, int ncols, const int8_t* q, int q0, int qstep, int qlen, int match, int mismatch,
                          int go, int ge, int lane) {
  Best3 best{0, 0x7fffffff, 0x7fffffff};
  const int n_strips = (qlen + 31) >> 5;
  for (int s = 0; s < n_strips; ++s) {
    const int i = 32 * s + lane + 1;
    const bool row_ok = i <= qlen;
    const int qc = row_ok ? q[q0 + (i - 1) * qstep] : 4;
    const short* hp = sm.hb[s & 1];
    const short* fp = sm.fb[s & 1];
    short* hc = sm.hb[(s + 1) & 1];
    short* fc = sm.fb[(s + 1) & 1];
    const bool last_row = lane == 31 || i == qlen;
    int e = 0, h_cur = 0, h_last = 0, f_out = 0;
    for (int t = 0; t < ncols + 31; ++t) {
      const int up_h = __shfl_up_sync(0xffffffffu, h_last, 1);
      const int up_f = __shfl_up_sync(0xffffffffu, f_out, 1);
      const int c = t - lane;
      if (c >= 0 && c < ncols && row_ok) {
        int diag, f_in;
        if (lane == 0) {
          diag = (s > 0 && c > 0) ? (int)hp[c - 1] : 0;
          f_in = s > 0 ? (int)fp[c] : 0;
        } else {
          diag = up_h;
          f_in = up_f;
        }
        const int rc = sm.ref[ref0 + c * rstep];
        const int sc = (rc == 4 || qc == 4) ? -mismatch : (rc == qc ? match : -mismatch);
        int hv = diag + sc;
        hv = max(hv, e);
        hv = max(hv, f_in);
        hv = max(hv, 0);
        e = max(0, max(hv - go, e - ge));
        f_out = max(0, max(hv - go, f_in - ge));
        h_last = h_cur;
        h_cur = hv;
        const Best3 cand{hv, c, i};
        if (better(cand, best)) best = cand;
        if (last_row) { hc[c] = (short)hv; fc[c] = (short)f_out; }
      }
    }
    __syncwarp();
  }
  // warp reduction of the lexicographic best
  for (int o = 16; o > 0; o >>= 1) {
    Best3 other;
    other.h = __shfl_xor_sync(0xffffffffu, best.h, o);
    other.col = __shfl_xor_sync(0xffffffffu, best.col, o);
    other.row = __shfl_xor_sync(0xffffffffu, best.row, o);
    if (better(other, best)) best = other;
  }
  return best;
}

__global__ void __launch_bounds__(32 * kSswWarps) dvb_ssw_scan_kernel(const int8_t* __restrict__ refs, const int8_t* __restrict__ queries,
                                                                     const SswJob* __restrict__ jobs, int n, int match, int mismatch, int go,
                                                                     int ge, SswEnds* __restrict__ out) {
  __shared__ WarpSmem smem[kSswWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a = blockIdx.x * kSswWarps + warp;
  if (a >= n) return;
  WarpSmem& sm = smem[warp];
  const SswJob job = jobs[a];
  SswEnds e{0, 0, 0, 0, 0, 0};
  if (job.ref_len > kSswMaxRef || job.q_len > 8000 || (long long)job.q_len * match > 30000) {   // int16 boundary rows / shared-memory budget
    e.status = 1;
    if (lane == 0) out[a] = e;
    return;
  }
  for (int i = lane; i < job.ref_len; i += 32) sm.ref[i] = refs[job.ref_off + i];
  __syncwarp();
  const int8_t* q = queries + job.q_off;
  const Best3 fwd = WarpScan(sm, 0, 1, job.ref_len, q, 0, 1, job.q_len, match, mismatch, go, ge, lane);
  e.score = fwd.h;
  if (fwd.h > 0) {
    e.ref_end = fwd.col;
    e.q_end = fwd.row - 1;
    __syncwarp();
    const Best3 rev = WarpScan(sm, e.ref_end, -1, e.ref_end + 1, q, e.q_end, -1, e.q_end + 1, match, mismatch, go, ge, lane);
    e.ref_begin = e.ref_end - rev.col;
    e.q_begin = e.q_end - (rev.row - 1);
    if (rev.h != fwd.h) e.status = 2;      // cannot happen (the reversed problem holds the forward optimum); the host path takes over if it does
  }
  if (lane == 0) out[a] = e;
}

}  // namespace

extern "C" {

// n alignments in one launch: host strings in, DvbSswAlignment + cigar strings out (cigars = char[n][cigar_stride], NUL-terminated;
// an alignment whose cigar does not fit gets cigar_len set and an empty string).  Results equal dvb_ssw_align's, field for field.
int dvb_ssw_align_batch(const char* const* refs, const int64_t* ref_lens, const char* const* queries, const int64_t* query_lens, int32_t n,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t device, DvbSswAlignment* out,
                        char* cigars, int64_t cigar_stride) {
  if (n < 0 || (n > 0 && (!refs || !ref_lens || !queries || !query_lens || !out)) || (cigar_stride > 0 && !cigars))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_ssw_align_batch: bad arguments");
  if (n == 0) return DVB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return dvb::fail(DVB_ERR_NO_DEVICE, "no CUDA device (dvb_ssw_align is the host entry point)");
  if (device < 0 || device >= ndev) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  DVB_CUDA(cudaSetDevice(device));
  std::vector<SswJob> jobs((size_t)n);
  long long rtot = 0, qtot = 0;
  for (int i = 0; i < n; ++i) {
    if (ref_lens[i] < 0 || query_lens[i] < 0 || ref_lens[i] > 0x7fffffff || query_lens[i] > 0x7fffffff)
      return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_ssw_align_batch: bad length");
    jobs[(size_t)i] = SswJob{rtot, qtot, (int)ref_lens[i], (int)query_lens[i]};
    rtot += ref_lens[i];
    qtot += query_lens[i];
  }
  std::vector<int8_t> rc((size_t)std::max<long long>(rtot, 1)), qc((size_t)std::max<long long>(qtot, 1));
  for (int i = 0; i < n; ++i) {
    for (int64_t k = 0; k < ref_lens[i]; ++k) rc[(size_t)(jobs[(size_t)i].ref_off + k)] = dvb_ssw_internal::BaseCode(refs[i][k]);
    for (int64_t k = 0; k < query_lens[i]; ++k) qc[(size_t)(jobs[(size_t)i].q_off + k)] = dvb_ssw_internal::BaseCode(queries[i][k]);
  }
  dvb::DevBuf d_r, d_q, d_jobs, d_out;
  auto release = [&]() { d_r.release(); d_q.release(); d_jobs.release(); d_out.release(); };
  cudaError_t ce = d_r.reserve(rc.size());
  if (ce == cudaSuccess) ce = d_q.reserve(qc.size());
  if (ce == cudaSuccess) ce = d_jobs.reserve(jobs.size() * sizeof(SswJob));
  if (ce == cudaSuccess) ce = d_out.reserve((size_t)n * sizeof(SswEnds));
  if (ce == cudaSuccess) ce = cudaMemcpy(d_r.p, rc.data(), rc.size(), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_q.p, qc.data(), qc.size(), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_jobs.p, jobs.data(), jobs.size() * sizeof(SswJob), cudaMemcpyHostToDevice);
  std::vector<SswEnds> ends((size_t)n);
  if (ce == cudaSuccess) {
    dvb_ssw_scan_kernel<<<(unsigned)((n + kSswWarps - 1) / kSswWarps), 32 * kSswWarps>>>(
        static_cast<const int8_t*>(d_r.p), static_cast<const int8_t*>(d_q.p), static_cast<const SswJob*>(d_jobs.p), n, match, mismatch, gap_open,
        gap_extend, static_cast<SswEnds*>(d_out.p));
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpy(ends.data(), d_out.p, (size_t)n * sizeof(SswEnds), cudaMemcpyDeviceToHost);
  release();
  if (ce != cudaSuccess) return dvb::fail(DVB_ERR_CUDA, "dvb_ssw_align_batch: %s", cudaGetErrorString(ce));
  int8_t mat[25];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? (int8_t)-mismatch : (i == j ? (int8_t)match : (int8_t)-mismatch);
  for (int i = 0; i < n; ++i) {
    char* cg = cigar_stride > 0 ? cigars + (size_t)i * cigar_stride : nullptr;
    memset(&out[i], 0, sizeof(out[i]));
    if (cg) cg[0] = 0;
    const SswEnds& e = ends[(size_t)i];
    if (ref_lens[i] <= 0 || query_lens[i] <= 0) continue;
    if (e.status != 0) {      // beyond the kernel's shared-memory budget: the host scans
      int st = dvb_ssw_align(refs[i], ref_lens[i], queries[i], query_lens[i], match, mismatch, gap_open, gap_extend, &out[i], cg, cigar_stride);
      if (st != DVB_OK) return st;
      continue;
    }
    out[i].sw_score = e.score;
    if (e.score <= 0) continue;
    out[i].ref_end = e.ref_end; out[i].query_end = e.q_end; out[i].ref_begin = e.ref_begin; out[i].query_begin = e.q_begin;
    int st = dvb_ssw_internal::FinishFromEnds(rc.data() + jobs[(size_t)i].ref_off, qc.data() + jobs[(size_t)i].q_off, (int)query_lens[i], mat,
                                              gap_open, gap_extend, &out[i], cg, cigar_stride);
    if (st != DVB_OK) return st;
  }
  return DVB_OK;
}

}  // extern "C"
