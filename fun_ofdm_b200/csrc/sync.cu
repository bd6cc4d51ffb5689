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

constexpr int DT = 512;          // samples per detector CTA
constexpr int DET_THREADS = 256;
constexpr double PLATEAU_THRESHOLD = 0.9; // frame_detector.h:12
constexpr double LTS_CORR_THRESHOLD = 0.9; // timing_sync.h:12

enum : uint8_t { TAG_NONE = 0, TAG_STS_START = 1, TAG_STS_END = 2, TAG_LTS1 = 4, TAG_LTS2 = 5 }; // tagged_vector.h:25-34

template <int FMT>
__global__ void __launch_bounds__(DET_THREADS) detect_kernel(const void *iq, double scale, uint64_t n, uint8_t *tags,
                                                             uint64_t *ev_x, uint32_t *ev_count, uint32_t ev_cap,
                                                             uint64_t x_limit)
{
    __shared__ double2 s_s[DT + 48];
    __shared__ double2 s_c[DT + 32];
    __shared__ double s_p[DT + 32];
    __shared__ uint8_t s_f[DT + 16];
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * DT;

    for (int k = tid; k < DT + 48; k += DET_THREADS) {
        const int64_t i = t0 - 48 + k;
        s_s[k] = (i >= 0 && (uint64_t)i < n) ? load_sample<FMT>(iq, (uint64_t)i, scale) : make_double2(0.0, 0.0); // the carry-over starts as zeros (frame_detector.cpp:27)
    }
    __syncthreads();
    // products of stream index t0 - 32 + j
    for (int j = tid; j < DT + 32; j += DET_THREADS) {
        const double2 a = s_s[j + 16], d = s_s[j];
        // input * std::conj(delayed): (a.x + i a.y)(d.x - i d.y)
        double cr = __dadd_rn(__dmul_rn(a.x, d.x), __dmul_rn(a.y, d.y));
        double ci = __dsub_rn(__dmul_rn(a.y, d.x), __dmul_rn(a.x, d.y));
        double pw = __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)); // std::norm
        if (cr != cr || ci != ci) { cr = 0.0; ci = 0.0; } // circular_accumulator.h:90
        if (pw != pw) pw = 0.0;
        s_c[j] = make_double2(cr, ci);
        s_p[j] = pw;
    }
    __syncthreads();
    // flag of stream index t0 - 16 + j: window = products j + 1 .. j + 16
    for (int j = tid; j < DT + 16; j += DET_THREADS) {
        double cr = 0.0, ci = 0.0, pw = 0.0;
#pragma unroll
        for (int k = 1; k <= 16; k++) {
            cr = __dadd_rn(cr, s_c[j + k].x);
            ci = __dadd_rn(ci, s_c[j + k].y);
            pw = __dadd_rn(pw, s_p[j + k]);
        }
        const double corr = hypot(cr, ci) / pw;
        s_f[j] = corr > PLATEAU_THRESHOLD ? 1 : 0;
    }
    __syncthreads();
    for (int m = tid; m < DT; m += DET_THREADS) {
        const int64_t i = t0 + m;
        if ((uint64_t)i >= n) break;
        int run = 0; // flags of i-15 .. i-1
#pragma unroll
        for (int k = 1; k <= 15; k++) run += s_f[m + k];
        const int f_old = s_f[m], f_now = s_f[m + 16];
        uint8_t tag = TAG_NONE;
        if (f_now && run == 15 && !f_old) tag = TAG_STS_START;  // plateau length reaches 16 exactly here
        else if (!f_now && run == 15 && f_old) tag = TAG_STS_END; // first clear flag after a run of >= 16
        if (tags) tags[i] = tag;
        if (tag == TAG_STS_END && (uint64_t)i < x_limit) {
            const uint32_t slot = atomicAdd(ev_count, 1u);
            if (slot < ev_cap) ev_x[slot] = (uint64_t)i;
        }
    }
}

// One CTA per STS_END event: 96 candidate offsets, 64 taps each (timing_sync.cpp:74-87), then the peak logic
// (timing_sync.cpp:89-118) on one thread.
template <int FMT>
__global__ void __launch_bounds__(128) lts_sync_kernel(const void *iq, double scale, uint64_t n, const uint64_t *ev_x,
                                                       const uint32_t *ev_count, uint32_t ev_cap, SyncRec *rec)
{
    __shared__ double2 s_s[160];
    __shared__ double s_val[96];
    const int tid = threadIdx.x;
    const uint32_t n_ev = min(*ev_count, ev_cap);
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
                    const int lts_offset = first - 32; // relative to x; never below the stream start for x >= 32
                    if ((int64_t)x + lts_offset >= -160) {
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

// Rank of every event by stream position (x values are distinct: one tag per sample): order[rank] = event.
__global__ void __launch_bounds__(256) rank_events_kernel(const SyncRec *rec, const uint32_t *ev_count, uint32_t ev_cap,
                                                          uint32_t *order)
{
    __shared__ uint64_t s_x[256];
    const uint32_t E = min(*ev_count, ev_cap);
    const uint32_t e = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x * 256 >= E) return;
    const uint64_t x = e < E ? rec[e].x : 0;
    uint32_t rank = 0;
    for (uint32_t base = 0; base < E; base += 256) {
        __syncthreads();
        s_x[threadIdx.x] = (base + threadIdx.x < E) ? rec[base + threadIdx.x].x : ~0ull;
        __syncthreads();
        const uint32_t m = min(256u, E - base);
        for (uint32_t k = 0; k < m; k++) rank += s_x[k] < x;
    }
    if (e < E) order[rank] = e;
}

// Events -> frames in stream order (one CTA).  Frame k starts at the LTS1 tag of the k-th successful event;
// it owns the samples up to the next LTS1 tag (fft_symbols.cpp:42-51 restarts there).
constexpr int BF_THREADS = 1024;

__global__ void __launch_bounds__(BF_THREADS) build_frames_kernel(const SyncRec *rec, const uint32_t *ev_count, uint32_t ev_cap,
                                                                  uint64_t n, double2 rot_in, uint32_t max_frames,
                                                                  uint32_t *order, uint64_t *lts1, uint32_t *avail, FrameRot *rot,
                                                                  double *phase, uint8_t *tags, SyncSummary *summary)
{
    __shared__ uint32_t s_part[BF_THREADS];
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x;
    const uint32_t n_all = *ev_count;
    const uint32_t E = min(n_all, ev_cap);
    // is sorted event k the start of a new frame?  (found, and not the same LTS1 as the previous found event)
    auto is_frame = [&](uint32_t k) -> bool {
        const SyncRec &r = rec[order[k]];
        if (!r.found || r.lts1 < 0 || (uint64_t)r.lts1 >= n) return false;
        for (int64_t q = (int64_t)k - 1; q >= 0; q--) {
            const SyncRec &pr = rec[order[q]];
            if (pr.found) return pr.lts1 != r.lts1;
        }
        return true;
    };
    const uint32_t per = (E + BF_THREADS - 1) / BF_THREADS;
    const uint32_t k0 = min(E, tid * per), k1 = min(E, k0 + per);
    uint32_t cnt = 0;
    for (uint32_t k = k0; k < k1; k++) cnt += is_frame(k) ? 1u : 0u;
    s_part[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < BF_THREADS; i++) { const uint32_t c = s_part[i]; s_part[i] = run; run += c; }
        s_total = run;
    }
    __syncthreads();
    uint32_t idx = s_part[tid];
    for (uint32_t k = k0; k < k1; k++) {
        if (!is_frame(k)) continue;
        const SyncRec &r = rec[order[k]];
        if (idx < max_frames) {
            lts1[idx] = (uint64_t)r.lts1;
            FrameRot fr;
            fr.rot_new = r.rot;
            fr.rot_old = rot_in;
            for (int64_t q = (int64_t)k - 1; q >= 0; q--) {
                const SyncRec &pr = rec[order[q]];
                if (pr.found) { fr.rot_old = pr.rot; break; }
            }
            fr.from = r.x;
            rot[idx] = fr;
            if (phase) phase[idx] = r.phase;
            // samples until the next frame's LTS1 (or the end of the stream)
            uint64_t end = n;
            for (uint32_t q = k + 1; q < E; q++) {
                const SyncRec &nr = rec[order[q]];
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
    __syncthreads();
    if (tid == 0) {
        summary->n_events = n_all;
        summary->n_frames = min(s_total, max_frames);
        summary->overflow = (n_all > ev_cap ? n_all - ev_cap : 0u) + (s_total > max_frames ? s_total - max_frames : 0u);
        summary->reserved = 0;
        // m_phase_acc after the stream: the last successful event's
        double ph = 0.0;
        int have = 0;
        for (int64_t q = (int64_t)E - 1; q >= 0; q--) {
            const SyncRec &pr = rec[order[q]];
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

cudaError_t launch_sync(const SyncArgs &a, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(a.ev_count, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    if (a.n_samples > 0) {
        const uint64_t blocks = (a.n_samples + DT - 1) / DT;
        const uint64_t x_limit = a.n_samples > 160 ? a.n_samples - 160 : 0; // timing_sync.cpp:68: x < input.size() - CARRYOVER_LENGTH
        switch (a.fmt) {
            case FMT_FC64:
                detect_kernel<FMT_FC64><<<(unsigned)blocks, DET_THREADS, 0, s>>>(a.iq, a.scale, a.n_samples, a.tags, a.ev_x, a.ev_count, a.ev_cap, x_limit);
                lts_sync_kernel<FMT_FC64><<<a.ev_cap, 128, 0, s>>>(a.iq, a.scale, a.n_samples, a.ev_x, a.ev_count, a.ev_cap, a.rec);
                break;
            case FMT_FC32:
                detect_kernel<FMT_FC32><<<(unsigned)blocks, DET_THREADS, 0, s>>>(a.iq, a.scale, a.n_samples, a.tags, a.ev_x, a.ev_count, a.ev_cap, x_limit);
                lts_sync_kernel<FMT_FC32><<<a.ev_cap, 128, 0, s>>>(a.iq, a.scale, a.n_samples, a.ev_x, a.ev_count, a.ev_cap, a.rec);
                break;
            case FMT_SC16:
                detect_kernel<FMT_SC16><<<(unsigned)blocks, DET_THREADS, 0, s>>>(a.iq, a.scale, a.n_samples, a.tags, a.ev_x, a.ev_count, a.ev_cap, x_limit);
                lts_sync_kernel<FMT_SC16><<<a.ev_cap, 128, 0, s>>>(a.iq, a.scale, a.n_samples, a.ev_x, a.ev_count, a.ev_cap, a.rec);
                break;
            default: return cudaErrorInvalidValue;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        rank_events_kernel<<<(a.ev_cap + 255) / 256, 256, 0, s>>>(a.rec, a.ev_count, a.ev_cap, a.order);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    build_frames_kernel<<<1, BF_THREADS, 0, s>>>(a.rec, a.ev_count, a.ev_cap, a.n_samples, a.rot_in, a.max_frames, a.order,
                                                 a.lts1, a.avail, a.rot, a.phase, a.tags, a.summary);
    return cudaGetLastError();
}

} // namespace b200rx
