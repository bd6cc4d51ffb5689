// Frame detection and timing synchronisation of a contiguous sample stream resident in HBM: the two
// blocks in front of the hot path.
//
//   frame_detector::work   (src/frame_detector.cpp:41-92)   -> detect_kernel
//   timing_sync::work      (src/timing_sync.cpp:51-139)     -> lts_sync_kernel + build_frames_kernel
//
// What the reference computes, per sample x of the stream:
//   c[x] = s[x] * conj(s[x-16]),  p[x] = |s[x]|^2,  C = sum of the last 16 c, P = sum of the last 16 p,
//   flag = |C| / P > 0.9; STS_START where a run of flags reaches length 16, STS_END at the first clear flag
//   after such a run.  At every STS_END x: correlate the next 96 offsets p in [x, x+96) with the 64 conjugated
//   LTS samples, keep offsets whose normalised correlation exceeds 0.9, sort them by (value, index) descending,
//   and if the best one has a partner exactly 64 samples away among the best five, tag LTS1 = min - 8 and
//   LTS2 = LTS1 + 64 and set the constant phase rotation applied to every later sample to
//   arg(s[min + 127] * conj(lts[63])) (the CFO loop of timing_sync.cpp:108-111 never runs).
//
// Deviations, all below the level that moves a tag (stated in DESIGN.md):
//   * the reference's moving sums are running sums (circular_accumulator.h:88-95: sum -= oldest, sum += newest)
//     whose rounding residue depends on the whole stream history; here every window is summed afresh, oldest
//     to newest.  The two differ by ~1e-16 relative to the window power, which can change a decision only
//     when |C|/P is within ~1e-12 of 0.9 - or in stretches of exact zeros, where the reference divides one
//     rounding residue by another and this code sees 0/0 (no plateau).
//   * the conjugated LTS samples are computed from the 802.11a definition in double precision; the
//     reference's table (preamble.h:432-497) is printed to 12 digits.
//   * libm: hypot / atan2 / cos / sin of the device instead of glibc's.
// The 64-tap correlations themselves are evaluated in the reference's order with unfused operations.
#include "rx_internal.cuh"

#include <math.h>

namespace b200rx {

namespace {

__constant__ double2 c_lts_conj[64];

constexpr int DET_THREADS = 128;
constexpr int DET_PER_THREAD = 8;                       // flags per thread
constexpr int DET_FLAGS = DET_THREADS * DET_PER_THREAD; // flags per CTA: stream indices t0 - 16 .. t0 + DT - 1
constexpr int DT = DET_FLAGS - 16;                      // tags per CTA
constexpr int DET_PRODUCTS = DET_FLAGS + 15;            // products of stream indices t0 - 31 .. t0 + DT - 1
constexpr int DET_SAMPLES = DET_PRODUCTS + 16;          // samples of stream indices t0 - 47 .. t0 + DT - 1
constexpr double PLATEAU_THRESHOLD = 0.9; // frame_detector.h:12
constexpr double LTS_CORR_THRESHOLD = 0.9; // timing_sync.h:12

enum : uint8_t { TAG_NONE = 0, TAG_STS_START = 1, TAG_STS_END = 2, TAG_LTS1 = 4, TAG_LTS2 = 5 }; // tagged_vector.h:25-34

// one padding slot per 8 entries: a thread's 23 consecutive window entries start 8 apart from its neighbour's, and a
// stride of 9 doubles keeps the 16 lanes of a half warp on distinct banks
__device__ __forceinline__ int pad8(int q) { return q + (q >> 3); }

// sum of the 16 products ending at each of 8 consecutive positions: w[j] = p[j] + ... + p[j + 15], j = 0..7, from 23
// inputs by doubling (pairs, fours, eights, sixteens): 66 additions instead of 120
__device__ __forceinline__ void window16(const double *src, double (&w)[DET_PER_THREAD])
{
    double a[23];
#pragma unroll
    for (int k = 0; k < 23; k++) a[k] = src[pad8(k) - 0];
#pragma unroll
    for (int k = 0; k < 22; k++) a[k] = __dadd_rn(a[k], a[k + 1]);
#pragma unroll
    for (int k = 0; k < 20; k++) a[k] = __dadd_rn(a[k], a[k + 2]);
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = __dadd_rn(a[k], a[k + 4]);
#pragma unroll
    for (int k = 0; k < DET_PER_THREAD; k++) w[k] = __dadd_rn(a[k], a[k + 8]);
}

// frame_detector::work (frame_detector.cpp:41-92).  Per CTA DT tags; per thread 8 plateau flags.
template <int FMT>
__global__ void __launch_bounds__(DET_THREADS) detect_kernel(const void *iq, double scale, uint64_t n, uint8_t *tags,
                                                             CtaEvents *cta_ev, uint8_t *cta_cnt, uint64_t x_limit,
                                                             const ScanParams *sp)
{
    if (sp) { // graph replay: the grid covers the lane's capacity, the capture ends where this call's samples end
        n = sp->n_samples;
        x_limit = sp->x_limit;
        if ((uint64_t)blockIdx.x * DT >= n) return;
    }
    __shared__ double s_re[DET_SAMPLES], s_im[DET_SAMPLES];
    __shared__ uint32_t s_nev;
    __shared__ uint16_t s_ev[CtaEvents::CAP];
    __shared__ double s_cr[DET_PRODUCTS + DET_PRODUCTS / 8 + 8], s_ci[DET_PRODUCTS + DET_PRODUCTS / 8 + 8],
        s_pw[DET_PRODUCTS + DET_PRODUCTS / 8 + 8];
    __shared__ uint32_t s_bits[DET_FLAGS / 32 + 1];
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * DT;

    {
        // all of a thread's loads are issued before the first use: one HBM round trip per tile, not nine
        constexpr int PER = (DET_SAMPLES + DET_THREADS - 1) / DET_THREADS;
        double2 v[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int k = tid + u * DET_THREADS;
            const int64_t i = t0 - 47 + k;
            // the carry-over starts as zeros (frame_detector.cpp:27)
            v[u] = (k < DET_SAMPLES && i >= 0 && (uint64_t)i < n) ? load_sample<FMT>(iq, (uint64_t)i, scale) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int k = tid + u * DET_THREADS;
            if (k < DET_SAMPLES) { s_re[k] = v[u].x; s_im[k] = v[u].y; }
        }
    }
    if (tid <= DET_FLAGS / 32) s_bits[tid] = 0;
    if (tid == 0) s_nev = 0;
    __syncthreads();
    // product q belongs to stream index t0 - 31 + q: input * conj(input delayed by 16), and the input's power
    for (int q = tid; q < DET_PRODUCTS; q += DET_THREADS) {
        const double ax = s_re[q + 16], ay = s_im[q + 16], dx = s_re[q], dy = s_im[q];
        double cr = __dadd_rn(__dmul_rn(ax, dx), __dmul_rn(ay, dy));
        double ci = __dsub_rn(__dmul_rn(ay, dx), __dmul_rn(ax, dy));
        double pw = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)); // std::norm
        if (cr != cr || ci != ci) { cr = 0.0; ci = 0.0; } // circular_accumulator.h:90
        if (pw != pw) pw = 0.0;
        const int o = pad8(q);
        s_cr[o] = cr;
        s_ci[o] = ci;
        s_pw[o] = pw;
    }
    __syncthreads();
    // flags of stream indices t0 - 16 + 8 * tid + (0..7): |C| / P > 0.9 as |C|^2 > (0.9 P)^2 (P >= 0; 0/0 is no plateau)
    {
        const int o = pad8(DET_PER_THREAD * tid); // multiple of 9: pad8(8 tid + k) = o + pad8(k)
        double wr[DET_PER_THREAD], wi[DET_PER_THREAD], wp[DET_PER_THREAD];
        window16(s_cr + o, wr);
        window16(s_ci + o, wi);
        window16(s_pw + o, wp);
        uint32_t mask = 0;
#pragma unroll
        for (int k = 0; k < DET_PER_THREAD; k++) {
            const double t = PLATEAU_THRESHOLD * wp[k];
            const double m2 = wr[k] * wr[k] + wi[k] * wi[k];
            if (m2 > t * t) mask |= 1u << k;
        }
        if (mask) atomicOr(&s_bits[tid >> 2], mask << (8 * (tid & 3)));
    }
    __syncthreads();
    // tag of stream index t0 + m from the flags of t0 + m - 16 .. t0 + m = flag bits m .. m + 16
    for (int m = tid; m < DT; m += DET_THREADS) {
        const int64_t i = t0 + m;
        if ((uint64_t)i >= n) break;
        const uint32_t w = __funnelshift_r(s_bits[m >> 5], s_bits[(m >> 5) + 1], m & 31) & 0x1FFFFu;
        uint8_t tag = TAG_NONE;
        if (w == 0x1FFFEu) tag = TAG_STS_START;     // 16 flags in a row ending here, none before: plateau length reaches 16
        else if (w == 0x0FFFFu) tag = TAG_STS_END;  // first clear flag after a run of >= 16
        if (tags) tags[i] = tag;
        if (tag == TAG_STS_END && (uint64_t)i < x_limit) {
            const uint32_t slot = atomicAdd(&s_nev, 1u);
            if (slot < CtaEvents::CAP) s_ev[slot] = (uint16_t)m;
        }
    }
    __syncthreads();
    // this CTA's STS_END tags in stream order (a tag needs 16 set flags in front of it, so at most DT / 17 per CTA; more
    // than CAP only in pathological streams - the excess is counted, not kept)
    if (tid == 0) {
        const uint32_t cnt = s_nev, kept = cnt < CtaEvents::CAP ? cnt : CtaEvents::CAP;
        cta_cnt[blockIdx.x] = (uint8_t)(cnt < 255 ? cnt : 255);
        if (cnt == 0) return; // the list of a CTA without events is never read
        CtaEvents out;
        out.count = (uint16_t)cnt;
        for (uint32_t a = 0; a < CtaEvents::CAP; a++) out.off[a] = 0xFFFF;
        for (uint32_t a = 0; a < kept; a++) { // insertion sort, kept is 0 or 1 almost always
            const uint16_t v = s_ev[a];
            uint32_t b = a;
            while (b > 0 && out.off[b - 1] > v) { out.off[b] = out.off[b - 1]; b--; }
            out.off[b] = v;
        }
        cta_ev[blockIdx.x] = out;
    }
}

// Per-CTA lists -> one list of STS_END positions in stream order (one CTA; exclusive scan of the per-CTA counts).
// ev_count[0] = events kept, ev_count[1] = events lost (beyond ev_cap, or beyond a detector CTA's list).
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) scan_events_kernel(const CtaEvents *cta_ev, const uint8_t *cta_cnt, uint32_t n_ctas,
                                                                   uint64_t *ev_x, uint32_t *ev_count, uint32_t ev_cap,
                                                                   const ScanParams *sp)
{
    if (sp) n_ctas = (uint32_t)((sp->n_samples + DT - 1) / DT);
    __shared__ uint32_t s_warp[32], s_lost[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_ctas + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t c0 = min(n_ctas, tid * per), c1 = min(n_ctas, c0 + per);
    uint32_t mine = 0, lost = 0;
    for (uint32_t c = c0; c < c1; c++) {
        const uint32_t cnt = cta_cnt[c];
        mine += min(cnt, (uint32_t)CtaEvents::CAP);
        lost += cnt > CtaEvents::CAP ? cnt - CtaEvents::CAP : 0u;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    lost = __reduce_add_sync(0xFFFFFFFFu, lost);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_lost[warp] = lost;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += v;
        }
        s_warp[lane] = wi - w; // exclusive offset of each warp
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, wi, 31);
        const uint32_t lost_all = __reduce_add_sync(0xFFFFFFFFu, s_lost[lane]);
        if (lane == 0) {
            ev_count[0] = min(total, ev_cap);
            ev_count[1] = lost_all + (total > ev_cap ? total - ev_cap : 0u);
        }
    }
    __syncthreads();
    uint32_t pos = s_warp[warp] + incl - mine;
    for (uint32_t c = c0; c < c1; c++) {
        if (cta_cnt[c] == 0) continue;
        const CtaEvents e = cta_ev[c];
        const uint32_t kept = min((uint32_t)e.count, (uint32_t)CtaEvents::CAP);
        for (uint32_t a = 0; a < kept; a++, pos++)
            if (pos < ev_cap) ev_x[pos] = (uint64_t)c * DT + e.off[a];
    }
}

struct OriginList { int64_t v[16]; uint32_t n; };

// One CTA per STS_END event: 96 candidate offsets, 64 taps each (timing_sync.cpp:74-87), then the peak logic
// (timing_sync.cpp:89-118) on one thread.
template <int FMT>
__global__ void __launch_bounds__(128) lts_sync_kernel(const void *iq, double scale, uint64_t n, const uint64_t *ev_x,
                                                       const uint32_t *ev_count, uint32_t ev_cap, SyncRec *rec,
                                                       const int64_t *origins, uint32_t n_origins, const OriginList inl,
                                                       const ScanParams *sp)
{
    if (sp) n = sp->n_samples;
    __shared__ double2 s_s[160];
    __shared__ double s_val[96];
    const int tid = threadIdx.x;
    const uint32_t n_ev = min(ev_count[0], ev_cap);
    for (uint32_t e = blockIdx.x; e < n_ev; e += gridDim.x) {
        const uint64_t x = ev_x[e];
        __syncthreads();
        for (int k = tid; k < 160; k += 128) s_s[k] = (x + k < n) ? load_sample<FMT>(iq, x + k, scale) : make_double2(0.0, 0.0);
        __syncthreads();
        if (tid < 96) {
            double cr = 0.0, ci = 0.0, pw = 0.0;
            for (int s = 0; s < 64; s++) {
                const double2 a = s_s[tid + s], l = c_lts_conj[s];
                const double pr = __dsub_rn(__dmul_rn(a.x, l.x), __dmul_rn(a.y, l.y));
                const double pi = __dadd_rn(__dmul_rn(a.x, l.y), __dmul_rn(a.y, l.x));
                cr = __dadd_rn(cr, pr);
                ci = __dadd_rn(ci, pi);
                pw = __dadd_rn(pw, __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)));
            }
            s_val[tid] = hypot(cr, ci) / pw;
        }
        __syncthreads();
        if (tid == 0) {
            // the five largest (value, index) pairs in descending order = front of the sorted + reversed vector
            double bv[5];
            int bp[5];
            int nb = 0, n_peaks = 0;
            for (int p = 0; p < 96; p++) {
                const double v = s_val[p];
                if (!(v > LTS_CORR_THRESHOLD)) continue;
                n_peaks++;
                int pos = nb; // insertion point: before every entry that is smaller, or equal with a smaller index
                while (pos > 0 && (bv[pos - 1] < v || (bv[pos - 1] == v && bp[pos - 1] < p))) pos--;
                if (pos >= 5) continue;
                const int last = nb < 5 ? nb : 4;
                for (int k = last; k > pos; k--) { bv[k] = bv[k - 1]; bp[k] = bp[k - 1]; }
                bv[pos] = v; bp[pos] = p;
                if (nb < 5) nb++;
            }
            SyncRec r;
            r.x = x;
            r.found = 0;
            r.n_peaks = (uint32_t)n_peaks;
            r.lts1 = 0;
            r.phase = 0.0;
            r.rot = make_double2(1.0, 0.0);
            // timing_sync.cpp:93-99: s runs over {0} only (jump = 5 > 3), t over the first five peaks
            for (int t = 0; t < nb; t++) {
                const int d = bp[0] - bp[t];
                if (d == 64 || d == -64) {
                    const int first = bp[0] < bp[t] ? bp[0] : bp[t];
                    const int lts_offset = first - 32; // relative to x
                    // timing_sync.cpp:102 `if(lts_offset < 0) break;` is evaluated in the coordinates of the work() buffer
                    // that examines the tag: 160 carried-over samples in front of the caller's chunk.  One capture = one
                    // buffer starting 160 samples before the capture; a caller that reproduces a chunked stream passes
                    // the buffer origins of its chunks (b200rx_set_receive_origins), the last one <= x applies.
                    int64_t origin = -160;
                    for (uint32_t o = 0; o < n_origins; o++) {
                        const int64_t v = origins[o];
                        if (v <= (int64_t)x) origin = v; else break;
                    }
                    for (uint32_t o = 0; o < inl.n; o++) { // the short list of a streaming call (kernel arguments)
                        const int64_t v = inl.v[o];
                        if (v <= (int64_t)x) origin = v; else break;
                    }
                    if (sp)
                        for (uint32_t o = 0; o < sp->n_origins; o++) { // ... or of a graph-replayed one (parameter block)
                            const int64_t v = sp->origins[o];
                            if (v <= (int64_t)x) origin = v; else break;
                        }
                    if ((int64_t)x + lts_offset >= origin) {
                        r.found = 1;
                        r.lts1 = (int64_t)x + lts_offset + 24;
                        // m_phase_acc = arg(input[lts_offset + 32 + 2 * 64 - 1] * LTS_TIME_DOMAIN_CONJ[63])
                        const double2 a = s_s[lts_offset + 159], l = c_lts_conj[63];
                        const double pr = __dsub_rn(__dmul_rn(a.x, l.x), __dmul_rn(a.y, l.y));
                        const double pi = __dadd_rn(__dmul_rn(a.x, l.y), __dmul_rn(a.y, l.x));
                        r.phase = atan2(pi, pr);
                        r.rot = make_double2(cos(r.phase), sin(r.phase));
                    }
                    break;
                }
            }
            rec[e] = r;
        }
    }
}

// Events -> frames in stream order (one CTA).  Frame k starts at the LTS1 tag of the k-th successful event;
// it owns the samples up to the next LTS1 tag (fft_symbols.cpp:42-51 restarts there).
constexpr int BF_THREADS = 1024;

__global__ void __launch_bounds__(BF_THREADS) build_frames_kernel(const SyncRec *rec, const uint32_t *ev_count, uint32_t ev_cap,
                                                                  uint64_t n, double2 rot_in, uint32_t max_frames,
                                                                  uint64_t *lts1, uint32_t *avail, FrameRot *rot,
                                                                  double *phase, uint8_t *tags, SyncSummary *summary,
                                                                  const ScanParams *sp)
{
    if (sp) { n = sp->n_samples; rot_in = sp->rot_in; }
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t E = min(ev_count[0], ev_cap); // events kept, in stream order
    const uint32_t n_lost = ev_count[1];
    // is event k the start of a new frame?  (found, and not the same LTS1 as the previous found event)
    auto is_frame = [&](uint32_t k) -> bool {
        const SyncRec &r = rec[k];
        if (!r.found || r.lts1 < 0 || (uint64_t)r.lts1 >= n) return false;
        for (int64_t q = (int64_t)k - 1; q >= 0; q--) {
            const SyncRec &pr = rec[q];
            if (pr.found) return pr.lts1 != r.lts1;
        }
        return true;
    };
    const uint32_t per = (E + BF_THREADS - 1) / BF_THREADS; // <= 8 for the usual ev_cap: one bit per event
    const uint32_t k0 = min(E, tid * per), k1 = min(E, k0 + per);
    uint32_t cnt = 0;
    uint64_t frame_bits = 0; // is_frame of events k0 .. k0 + 63 (per > 64 falls back to re-evaluation)
    for (uint32_t k = k0; k < k1; k++) {
        const bool f = is_frame(k);
        cnt += f ? 1u : 0u;
        if (f && k - k0 < 64) frame_bits |= 1ull << (k - k0);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += v;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    uint32_t idx = s_warp[warp] + incl - cnt;
    for (uint32_t k = k0; k < k1; k++) {
        const bool f = (k - k0 < 64) ? ((frame_bits >> (k - k0)) & 1ull) != 0 : is_frame(k);
        if (!f) continue;
        const SyncRec &r = rec[k];
        if (idx < max_frames) {
            lts1[idx] = (uint64_t)r.lts1;
            FrameRot fr;
            fr.rot_new = r.rot;
            fr.rot_old = rot_in;
            for (int64_t q = (int64_t)k - 1; q >= 0; q--) {
                const SyncRec &pr = rec[q];
                if (pr.found) { fr.rot_old = pr.rot; break; }
            }
            fr.from = r.x;
            rot[idx] = fr;
            if (phase) phase[idx] = r.phase;
            // samples until the next frame's LTS1 (or the end of the stream)
            uint64_t end = n;
            for (uint32_t q = k + 1; q < E; q++) {
                const SyncRec &nr = rec[q];
                if (nr.found && nr.lts1 != r.lts1) { if (nr.lts1 > r.lts1 && (uint64_t)nr.lts1 < n) end = (uint64_t)nr.lts1; break; }
            }
            const uint64_t span = end - (uint64_t)r.lts1;
            avail[idx] = span > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)span;
        }
        if (tags) {
            tags[r.lts1] = TAG_LTS1;
            if ((uint64_t)r.lts1 + 64 < n) tags[r.lts1 + 64] = TAG_LTS2;
        }
        idx++;
    }
    if (tid == 0) {
        summary->n_events = E + n_lost;
        summary->n_frames = min(s_total, max_frames);
        summary->overflow = n_lost + (s_total > max_frames ? s_total - max_frames : 0u);
        summary->reserved = 0;
        // m_phase_acc after the stream: the last successful event's
        double ph = 0.0;
        int have = 0;
        for (int64_t q = (int64_t)E - 1; q >= 0; q--) {
            const SyncRec &pr = rec[q];
            if (pr.found) { ph = pr.phase; have = 1; break; }
        }
        summary->last_phase = ph;
        summary->phase_valid = (uint32_t)have;
        summary->pad = 0;
    }
}

} // namespace

cudaError_t upload_sync_tables()
{
    // Long training symbol in time (802.11a 17.3.3): x[n] = (1/64) sum_k L_k exp(+2 pi i k n / 64), k = -26..26
    static const char *lts = "++--++-+-++++++--++-+-++++0+--++-+-+-----++--+-+-++++";
    double2 tab[64];
    for (int nn = 0; nn < 64; nn++) {
        double re = 0.0, im = 0.0;
        for (int k = -26; k <= 26; k++) {
            const double l = lts[k + 26] == '+' ? 1.0 : (lts[k + 26] == '-' ? -1.0 : 0.0);
            const int ph = ((k * nn) % 64 + 64) % 64;
            re += l * cos(2.0 * M_PI * ph / 64.0);
            im += l * sin(2.0 * M_PI * ph / 64.0);
        }
        tab[nn] = make_double2(re / 64.0, -im / 64.0); // conjugate
    }
    return cudaMemcpyToSymbol(c_lts_conj, tab, sizeof(tab));
}

uint32_t sync_cta_count(uint64_t n_samples) { return (uint32_t)((n_samples + DT - 1) / DT); }

cudaError_t launch_sync(const SyncArgs &a, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(a.ev_count, 0, 2 * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    const uint64_t grid_n = a.sp ? a.grid_samples : a.n_samples; // graph replay: launch for the lane's capacity
    if (grid_n > 0) {
        const unsigned blocks = sync_cta_count(grid_n);
        const uint64_t x_limit = a.n_samples > 160 ? a.n_samples - 160 : 0; // timing_sync.cpp:68: x < input.size() - CARRYOVER_LENGTH
        switch (a.fmt) {
            case FMT_FC64: detect_kernel<FMT_FC64><<<blocks, DET_THREADS, 0, s>>>(a.iq, a.scale, a.n_samples, a.tags, a.cta_ev, a.cta_cnt, x_limit, a.sp); break;
            case FMT_FC32: detect_kernel<FMT_FC32><<<blocks, DET_THREADS, 0, s>>>(a.iq, a.scale, a.n_samples, a.tags, a.cta_ev, a.cta_cnt, x_limit, a.sp); break;
            case FMT_SC16: detect_kernel<FMT_SC16><<<blocks, DET_THREADS, 0, s>>>(a.iq, a.scale, a.n_samples, a.tags, a.cta_ev, a.cta_cnt, x_limit, a.sp); break;
            default: return cudaErrorInvalidValue;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        scan_events_kernel<<<1, SCAN_THREADS, 0, s>>>(a.cta_ev, a.cta_cnt, blocks, a.ev_x, a.ev_count, a.ev_cap, a.sp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        // enough CTAs for one event each at the usual rates; the kernel strides over the rest
        // (an STS_END tag needs 17 samples of its own: a capture of n samples holds at most n / 17 events)
        unsigned lts_grid = a.ev_cap < 4096u ? a.ev_cap : 4096u;
        const uint64_t ev_bound = grid_n / 17 + 1;
        if (ev_bound < lts_grid) lts_grid = (unsigned)ev_bound;
        OriginList inl;
        inl.n = a.n_inline <= 16 ? a.n_inline : 0;
        for (uint32_t o = 0; o < 16; o++) inl.v[o] = o < inl.n ? a.origins_inline[o] : 0;
        switch (a.fmt) {
            case FMT_FC64: lts_sync_kernel<FMT_FC64><<<lts_grid, 128, 0, s>>>(a.iq, a.scale, a.n_samples, a.ev_x, a.ev_count, a.ev_cap, a.rec, a.origins, a.n_origins, inl, a.sp); break;
            case FMT_FC32: lts_sync_kernel<FMT_FC32><<<lts_grid, 128, 0, s>>>(a.iq, a.scale, a.n_samples, a.ev_x, a.ev_count, a.ev_cap, a.rec, a.origins, a.n_origins, inl, a.sp); break;
            default: lts_sync_kernel<FMT_SC16><<<lts_grid, 128, 0, s>>>(a.iq, a.scale, a.n_samples, a.ev_x, a.ev_count, a.ev_cap, a.rec, a.origins, a.n_origins, inl, a.sp); break;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    build_frames_kernel<<<1, BF_THREADS, 0, s>>>(a.rec, a.ev_count, a.ev_cap, a.n_samples, a.rot_in, a.max_frames,
                                                 a.lts1, a.avail, a.rot, a.phase, a.tags, a.summary, a.sp);
    return cudaGetLastError();
}

} // namespace b200rx
