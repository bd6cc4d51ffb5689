// C ABI of libb200rx.so (include/b200rx.h): handle, device scratch, stream plumbing, stage timing.
// No CPU implementation of any stage exists in this library: without a CUDA device every entry
// point fails.
#include "rx_internal.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace b200rx;

static thread_local char g_create_error[512] = {0};

struct b200rx_handle {
    int device = 0;
    b200rx_limits limits{};
    uint32_t max_steps = 0;  // trellis capacity per frame (multiple of 32)
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    // optional per-call stage timing over many calls (b200rx_profile_begin/read)
    std::vector<cudaEvent_t> ring; // 4 events per slot
    uint32_t ring_slots = 0, ring_used = 0;

    // device scratch of the call being issued (one of the sets below)
    FrameDesc *desc = nullptr;
    uint32_t *bm = nullptr;  // front end -> ACS: metric words (4 B per step) or soft-symbol pairs (2 B per step)
    uint32_t *dec = nullptr; // survivor words: 2 per trellis step per frame
    unsigned long long *counters = nullptr; // 8 words (5 used)
    double2 *hinv = nullptr; // split front end: inverse channel per frame, header pass -> data kernel
    Tuning tn;

    // staging of the host-buffer entry points (grow-only).  b200rx_submit_batch keeps up to B200RX_MAX_INFLIGHT calls in
    // flight, each on its own slot: staging + decode scratch + completion event.  Slot 0's scratch is lane 0's.
    struct HostSlot {
        uint8_t *d_iq = nullptr; size_t d_iq_cap = 0; // bytes, samples in the handle's format
        uint64_t *d_lts1 = nullptr;
        uint32_t *d_avail = nullptr;
        uint8_t *d_payload = nullptr; size_t d_payload_cap = 0;
        uint16_t *d_len = nullptr;
        uint8_t *d_rate = nullptr;
        uint8_t *d_status = nullptr;
        FrameDesc *desc = nullptr; uint32_t *bm = nullptr; uint32_t *dec = nullptr; unsigned long long *counters = nullptr;
        double2 *hinv = nullptr;
        std::vector<cudaEvent_t> pipe_ev; // 2 per chunk (samples landed, results ready) + 1
        cudaEvent_t done = nullptr;
        bool busy = false;
        uint64_t ticket = 0;
    };
    HostSlot hs[B200RX_MAX_INFLIGHT];
    uint64_t next_ticket = 1;

    // pipeline depth > 1: consecutive device-buffer calls rotate over `depth` lanes (own scratch set, own stream),
    // so that batch j+1 starts while batch j is still in its Viterbi kernel (b200rx_set_pipeline_depth)
    struct Lane {
        FrameDesc *desc = nullptr; uint32_t *bm = nullptr; uint32_t *dec = nullptr; unsigned long long *counters = nullptr;
        double2 *hinv = nullptr;
        cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; bool used = false;
    };
    Lane lanes[B200RX_MAX_PIPELINE_DEPTH];
    uint32_t depth = 1;
    uint64_t call_idx = 0;
    cudaEvent_t ev_in = nullptr;

    // host-buffer pipeline: H2D of chunk i+1 overlaps the kernels of chunk i
    cudaStream_t copy_stream = nullptr, pull_stream = nullptr, aux_stream[7] = {}, d2h_stream = nullptr;

    // frame detection / timing synchronisation scratch, one set per pipeline lane (allocated on first use)
    struct SyncScratch {
        uint64_t *ev_x = nullptr;
        uint32_t *ev_count = nullptr;
        SyncRec *rec = nullptr;
        CtaEvents *cta_ev = nullptr; uint8_t *cta_cnt = nullptr; uint32_t cta_cap = 0;
        uint64_t *lts1 = nullptr;
        uint32_t *avail = nullptr;
        FrameRot *rot = nullptr;
        double *phase = nullptr;
        SyncSummary *summary = nullptr; // device
    };
    SyncScratch sy[B200RX_MAX_PIPELINE_DEPTH];
    uint32_t sy_ev_cap = 0;
    SyncSummary *sy_summary_host = nullptr; // pinned

    // work() buffer origins for the next raw-capture call (b200rx_set_receive_origins); device copy per sync lane
    std::vector<int64_t> origins;
    int64_t *d_origins[B200RX_MAX_PIPELINE_DEPTH] = {};
    uint32_t d_origins_cap[B200RX_MAX_PIPELINE_DEPTH] = {};

    // two-phase passes (b200rx_pass_*): per lane a stream, sample / output staging and the host-mapped frame list
    struct PassLane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        uint8_t *d_iq = nullptr; size_t d_iq_cap = 0;          // bytes, samples in the handle's format
        uint8_t *d_payload = nullptr; size_t d_payload_cap = 0;
        uint8_t *d_status = nullptr, *d_select = nullptr;      // [max_frames]
        uint8_t *h_list = nullptr, *h_list_dev = nullptr;      // pinned + mapped: SyncSummary, then b200rx_pass_frame[max_frames]
        uint8_t *h_select = nullptr;                           // pinned [max_frames]
        // the scan phase as a CUDA graph (captured once per lane; re-captured when the staging moves or grows)
        cudaGraphExec_t scan_exec = nullptr;
        ScanParams *h_sp = nullptr, *d_sp = nullptr;           // pinned / device parameter block
        const uint8_t *scan_iq = nullptr; uint64_t scan_cap = 0; int scan_fmt = -1; double scan_scale = 0.0;
        uint64_t fill = 0;                                     // samples staged so far
        uint32_t n_frames = 0;
        bool scanned = false, busy = false, tagged = false;
        uint64_t ticket = 0;
    };
    PassLane pl[B200RX_MAX_PIPELINE_DEPTH];
    int pass_lane = -1;
    uint64_t pass_count = 0, pass_next_ticket = 1;

    // format of every `iq` argument (b200rx_set_sample_format)
    int fmt = FMT_FC64;
    double scale = 1.0;
    int sm_count = 148;

    uint64_t launches = 0;
    char error_buf[512] = {0}; // what b200rx_last_error returns (no allocation on the error path)
};

namespace {

int fail(b200rx_handle *h, int code, const char *what, cudaError_t ce = cudaSuccess)
{
    char *buf = h ? h->error_buf : g_create_error;
    if (ce != cudaSuccess) snprintf(buf, 512, "%s: %s (%s)", what, cudaGetErrorString(ce), cudaGetErrorName(ce));
    else snprintf(buf, 512, "%s", what);
    return code;
}

// events of the current call: a ring slot while profiling, the handle's single set otherwise
inline cudaEvent_t *call_events(b200rx_handle *h)
{
    if (h->ring_slots && h->ring_used < h->ring_slots) return &h->ring[4 * (size_t)(h->ring_used++)];
    return h->ev;
}

// scratch set 0 is the one every non-pipelined entry point uses
inline void use_lane(b200rx_handle *h, int i)
{
    h->desc = h->lanes[i].desc; h->bm = h->lanes[i].bm; h->dec = h->lanes[i].dec; h->counters = h->lanes[i].counters;
    h->hinv = h->lanes[i].hinv;
}

#define CU(h, call)                                                         \
    do {                                                                    \
        cudaError_t ce_ = (call);                                           \
        if (ce_ != cudaSuccess) return fail(h, B200RX_E_CUDA, #call, ce_);  \
    } while (0)

// Host-buffer calls still in flight (b200rx_submit_batch) own scratch the other entry points use: finish them first.
inline int quiesce_host(b200rx_handle *h)
{
    for (auto &S : h->hs)
        if (S.busy) {
            CU(h, cudaEventSynchronize(S.done));
            S.busy = false;
        }
    if (h->depth <= 1) use_lane(h, 0);
    return B200RX_OK;
}

// ---- constant tables (regenerated from their defining rules; pinned by tests against the reference) ----
std::mutex g_device_init_mutex;
bool g_device_ready[64] = {false}; // constant tables uploaded and function attributes set, per device

// one scratch set {desc, bm, dec, counters}: all four or none
struct ScratchPtrs { FrameDesc *desc; uint32_t *bm; uint32_t *dec; unsigned long long *counters; double2 *hinv; };

void free_scratch(ScratchPtrs &p)
{
    cudaFree(p.desc); cudaFree(p.bm); cudaFree(p.dec); cudaFree(p.counters); cudaFree(p.hinv);
    p.desc = nullptr; p.bm = nullptr; p.dec = nullptr; p.counters = nullptr; p.hinv = nullptr;
}

cudaError_t alloc_scratch(const b200rx_handle *h, ScratchPtrs &p)
{
    const size_t nf = h->limits.max_frames;
    p.desc = nullptr; p.bm = nullptr; p.dec = nullptr; p.counters = nullptr; p.hinv = nullptr;
    cudaError_t e = cudaMalloc((void **)&p.desc, nf * sizeof(FrameDesc));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p.bm, nf * (size_t)h->max_steps * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p.dec, nf * (size_t)h->max_steps * 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p.counters, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p.hinv, nf * 64 * sizeof(double2));
    if (e != cudaSuccess) {
        free_scratch(p);
        (void)cudaGetLastError();
    }
    return e;
}

} // namespace

namespace b200rx {
// defined next to the kernels that own the __constant__ symbols
cudaError_t upload_frontend_tables(const double2 *tw, const int8_t *pol);
cudaError_t upload_viterbi_tables(const uint32_t *crc, const uint8_t *scr, const uint32_t *xpow);
}

cudaError_t b200rx::upload_tables()
{
    // twiddles exp(-2 pi i k / 64), built with exact octant symmetry
    double2 tw[64];
    for (int k = 0; k < 64; k++) {
        const int q = k % 16, quad = k / 16;
        double c, s;
        if (q == 0) { c = 1.0; s = 0.0; }
        else if (q == 8) { c = s = sqrt(0.5); }
        else if (q < 8) { c = cos(2.0 * M_PI * q / 64.0); s = sin(2.0 * M_PI * q / 64.0); }
        else { c = sin(2.0 * M_PI * (16 - q) / 64.0); s = cos(2.0 * M_PI * (16 - q) / 64.0); }
        double re, im; // exp(+i theta)
        switch (quad) {
            case 0: re = c; im = s; break;
            case 1: re = -s; im = c; break;
            case 2: re = -c; im = -s; break;
            default: re = s; im = -c; break;
        }
        tw[k] = make_double2(re, -im);
    }
    // pilot polarity: 802.11a 17.3.5.9, scrambler x^7 + x^4 + 1 from the all-ones state, 0 -> +1, 1 -> -1
    int8_t pol[127];
    int st = 0x7F;
    for (int i = 0; i < 127; i++) {
        const int fb = ((st >> 6) ^ (st >> 3)) & 1;
        st = ((st << 1) | fb) & 0x7F;
        pol[i] = fb ? -1 : 1;
    }
    cudaError_t e = upload_frontend_tables(tw, pol);
    if (e != cudaSuccess) return e;
    e = upload_sync_tables();
    if (e != cudaSuccess) return e;

    // CRC-32/ISO-HDLC slice-by-4 tables
    static uint32_t crc[4][256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        crc[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 4; t++) crc[t][i] = (crc[t - 1][i] >> 8) ^ crc[0][crc[t - 1][i] & 0xFF];
    // descrambler: ppdu.cpp:257-263, state 93, feedback = bit6 ^ bit3, one step per byte
    uint8_t scr[127];
    int state = 93;
    for (int x = 0; x < 127; x++) {
        const int fb = ((state >> 6) & 1) ^ ((state >> 3) & 1);
        scr[x] = (uint8_t)fb;
        state = ((state << 1) & 0x7E) | fb;
    }
    // CRC combine factors x^(8 * L * 2^lvl) mod P (reflected representation, bit 31 = x^0), for the
    // warp-parallel CRC of the traceback kernel
    auto multmodp = [](uint32_t a, uint32_t b) {
        uint32_t p = 0;
        for (int i = 31; i >= 0; i--) {
            if ((a >> i) & 1u) p ^= b;
            b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
        }
        return p;
    };
    static uint32_t xpow[5][132];
    const uint32_t x8 = 1u << 23; // x^8
    uint32_t cur = 1u << 31;      // x^0
    for (int L = 0; L < 132; L++) {
        uint32_t f = cur;
        for (int lvl = 0; lvl < 5; lvl++) { xpow[lvl][L] = f; f = multmodp(f, f); }
        cur = multmodp(cur, x8);
    }
    return upload_viterbi_tables(&crc[0][0], scr, &xpow[0][0]);
}

extern "C" {

const char *b200rx_version(void) { return "b200rx 0.1 (sm_100a)"; }

const char *b200rx_last_error(const b200rx_handle *h) { return h ? h->error_buf : g_create_error; }

int b200rx_create(int device, const b200rx_limits *limits, b200rx_handle **out)
{
    if (!out || !limits) return fail(nullptr, B200RX_E_ARG, "b200rx_create: null argument");
    *out = nullptr;
    if (limits->max_frames == 0 || limits->max_payload_bytes > 4095)
        return fail(nullptr, B200RX_E_ARG, "b200rx_create: limits out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, B200RX_E_DEVICE, "b200rx_create: no CUDA device (this library has no CPU path)", ce);
    if (device < 0 || device >= ndev) return fail(nullptr, B200RX_E_ARG, "b200rx_create: bad device index");
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char buf[200];
        snprintf(buf, sizeof(buf), "b200rx_create: device %d is sm_%d%d; the kernels are built for sm_100a only", device,
                 prop.major, prop.minor);
        return fail(nullptr, B200RX_E_DEVICE, buf);
    }
    CU(nullptr, cudaSetDevice(device));
    {
        std::lock_guard<std::mutex> lock(g_device_init_mutex);
        if (device >= 64 || !g_device_ready[device]) {
            CU(nullptr, upload_tables());
            CU(nullptr, prepare_device_functions());
            if (device < 64) g_device_ready[device] = true;
        }
    }

    b200rx_handle *h = new (std::nothrow) b200rx_handle();
    if (!h) return fail(nullptr, B200RX_E_NOMEM, "b200rx_create: out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->tn.sm_count = prop.multiProcessorCount;
    h->limits = *limits;
    h->max_steps = max_steps_for(limits->max_payload_bytes);
    const size_t nf = limits->max_frames;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->pull_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 7; i++)
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->aux_stream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&h->ev[i]);
    if (e == cudaSuccess) { // scratch set 0: lane 0 and host slot 0 share it
        ScratchPtrs sp;
        e = alloc_scratch(h, sp);
        h->lanes[0].desc = sp.desc; h->lanes[0].bm = sp.bm; h->lanes[0].dec = sp.dec; h->lanes[0].counters = sp.counters;
        h->lanes[0].hinv = sp.hinv;
        h->hs[0].desc = sp.desc; h->hs[0].bm = sp.bm; h->hs[0].dec = sp.dec; h->hs[0].counters = sp.counters;
        h->hs[0].hinv = sp.hinv;
        use_lane(h, 0);
    }
    A((void **)&h->hs[0].d_lts1, nf * sizeof(uint64_t));
    A((void **)&h->hs[0].d_avail, nf * sizeof(uint32_t));
    A((void **)&h->hs[0].d_len, nf * sizeof(uint16_t));
    A((void **)&h->hs[0].d_rate, nf);
    A((void **)&h->hs[0].d_status, nf);
    if (e != cudaSuccess) {
        int code = (e == cudaErrorMemoryAllocation) ? B200RX_E_NOMEM : B200RX_E_CUDA;
        fail(nullptr, code, "b200rx_create: allocating device scratch", e);
        (void)cudaGetLastError();
        b200rx_destroy(h);
        return code;
    }
    h->stream = h->own_stream;
    *out = h;
    return B200RX_OK;
}

int b200rx_destroy(b200rx_handle *h)
{
    if (!h) return B200RX_OK;
    cudaSetDevice(h->device);
    if (h->stream && h->lanes[0].desc) cudaStreamSynchronize(h->stream);
    for (int i = 0; i < B200RX_MAX_PIPELINE_DEPTH; i++) {
        b200rx_handle::Lane &l = h->lanes[i];
        if (l.stream) { cudaStreamSynchronize(l.stream); cudaStreamDestroy(l.stream); }
        if (l.done) cudaEventDestroy(l.done);
        cudaFree(l.desc); cudaFree(l.bm); cudaFree(l.dec); cudaFree(l.counters); cudaFree(l.hinv); // set 0 included (shared with host slot 0)
    }
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    for (auto &y : h->sy) {
        cudaFree(y.ev_x); cudaFree(y.ev_count); cudaFree(y.rec); cudaFree(y.cta_ev); cudaFree(y.cta_cnt); cudaFree(y.lts1);
        cudaFree(y.avail); cudaFree(y.rot); cudaFree(y.phase); cudaFree(y.summary);
    }
    for (auto p : h->d_origins) cudaFree(p);
    if (h->sy_summary_host) cudaFreeHost(h->sy_summary_host);
    for (int i = 0; i < B200RX_MAX_INFLIGHT; i++) {
        b200rx_handle::HostSlot &S = h->hs[i];
        if (S.done) { cudaEventSynchronize(S.done); cudaEventDestroy(S.done); }
        cudaFree(S.d_iq); cudaFree(S.d_lts1); cudaFree(S.d_avail); cudaFree(S.d_payload);
        cudaFree(S.d_len); cudaFree(S.d_rate); cudaFree(S.d_status);
        if (i > 0) { cudaFree(S.desc); cudaFree(S.bm); cudaFree(S.dec); cudaFree(S.counters); cudaFree(S.hinv); }
        for (cudaEvent_t e : S.pipe_ev) if (e) cudaEventDestroy(e);
    }
    for (auto &P : h->pl) {
        if (P.stream) { cudaStreamSynchronize(P.stream); cudaStreamDestroy(P.stream); }
        if (P.done) cudaEventDestroy(P.done);
        cudaFree(P.d_iq); cudaFree(P.d_payload); cudaFree(P.d_status); cudaFree(P.d_select);
        if (P.h_list) cudaFreeHost(P.h_list);
        if (P.h_select) cudaFreeHost(P.h_select);
        if (P.scan_exec) cudaGraphExecDestroy(P.scan_exec);
        if (P.h_sp) cudaFreeHost(P.h_sp);
        cudaFree(P.d_sp);
    }
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (cudaEvent_t e : h->ring) if (e) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->pull_stream) cudaStreamDestroy(h->pull_stream);
    for (int i = 0; i < 7; i++)
        if (h->aux_stream[i]) cudaStreamDestroy(h->aux_stream[i]);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    (void)cudaGetLastError();
    delete h;
    return B200RX_OK;
}

int b200rx_set_stream(b200rx_handle *h, void *cuda_stream)
{
    if (!h) return B200RX_E_ARG;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return B200RX_OK;
}

int b200rx_set_sample_format(b200rx_handle *h, int format, double sc16_scale)
{
    if (!h) return B200RX_E_ARG;
    if (format != B200RX_FMT_FC64 && format != B200RX_FMT_FC32 && format != B200RX_FMT_SC16 && format != B200RX_FMT_TAGGED_FC64)
        return fail(h, B200RX_E_ARG, "b200rx_set_sample_format: unknown format");
    if (format == B200RX_FMT_SC16 && !(sc16_scale > 0.0)) return fail(h, B200RX_E_ARG, "b200rx_set_sample_format: scale must be > 0");
    h->fmt = format;
    h->scale = format == B200RX_FMT_SC16 ? sc16_scale : 1.0;
    return B200RX_OK;
}

int b200rx_set_tuning(b200rx_handle *h, const char *key, int64_t value)
{
    if (!h || !key) return B200RX_E_ARG;
    const int v = (int)value;
    if (!strcmp(key, "acs_gen")) { if (v != 2 && v != 3) return fail(h, B200RX_E_ARG, "b200rx_set_tuning: acs_gen is 2 or 3"); }
    int *field = nullptr;
    if (!strcmp(key, "acs_gen")) field = &h->tn.acs_gen;
    else if (!strcmp(key, "acs_lb")) field = &h->tn.acs_lb;
    else if (!strcmp(key, "acs_warps")) field = &h->tn.acs_warps;
    else if (!strcmp(key, "acs_rn")) field = &h->tn.acs_rn;
    else if (!strcmp(key, "h2d_chunk")) field = &h->tn.h2d_chunk;
    else if (!strcmp(key, "h2d_chunk_min")) field = &h->tn.h2d_chunk_min;
    else if (!strcmp(key, "pull_mode")) field = &h->tn.pull_mode;
    else if (!strcmp(key, "fe_split")) field = &h->tn.fe_split;
    else if (!strcmp(key, "scan_graph")) field = &h->tn.scan_graph;
    else return fail(h, B200RX_E_ARG, "b200rx_set_tuning: unknown key");
    int rc = b200rx_synchronize(h); // nothing in flight may see two settings
    if (rc != B200RX_OK) return rc;
    *field = v;
    return B200RX_OK;
}

int b200rx_set_receive_origins(b200rx_handle *h, const int64_t *origins, uint32_t n)
try {
    if (!h || (n && !origins)) return B200RX_E_ARG;
    for (uint32_t i = 1; i < n; i++)
        if (origins[i] < origins[i - 1]) return fail(h, B200RX_E_ARG, "b200rx_set_receive_origins: origins must ascend");
    h->origins.assign(origins, origins + n);
    return B200RX_OK;
}
catch (...) { return B200RX_E_NOMEM; } // std::bad_alloc must not cross the C boundary

int b200rx_synchronize(b200rx_handle *h)
{
    if (!h) return B200RX_E_ARG;
    CU(h, cudaSetDevice(h->device));
    for (int i = 0; i < B200RX_MAX_PIPELINE_DEPTH; i++)
        if (h->lanes[i].stream && h->lanes[i].used) CU(h, cudaStreamSynchronize(h->lanes[i].stream));
    for (auto &S : h->hs)
        if (S.busy) {
            CU(h, cudaEventSynchronize(S.done));
            S.busy = false;
        }
    for (auto &P : h->pl)
        if (P.stream) {
            CU(h, cudaStreamSynchronize(P.stream));
            P.busy = false;
        }
    CU(h, cudaStreamSynchronize(h->stream));
    return B200RX_OK;
}

int b200rx_set_pipeline_depth(b200rx_handle *h, uint32_t depth)
{
    if (!h || depth < 1 || depth > B200RX_MAX_PIPELINE_DEPTH) return fail(h, B200RX_E_ARG, "b200rx_set_pipeline_depth: depth must be 1 .. B200RX_MAX_PIPELINE_DEPTH");
    int rc = b200rx_synchronize(h);
    if (rc != B200RX_OK) return rc;
    for (uint32_t i = 0; i < depth; i++) {
        b200rx_handle::Lane &l = h->lanes[i];
        if (depth > 1 && !l.stream) {
            CU(h, cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
            CU(h, cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        }
        if (i > 0 && !l.desc) { // a lane has its whole scratch set or none of it; the depth in force stays as it was on failure
            ScratchPtrs sp;
            cudaError_t e = alloc_scratch(h, sp);
            if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_set_pipeline_depth: scratch for an extra lane", e);
            l.desc = sp.desc; l.bm = sp.bm; l.dec = sp.dec; l.counters = sp.counters; l.hinv = sp.hinv;
        }
    }
    if (!h->ev_in) CU(h, cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    h->depth = depth;
    h->tn.inflight = (int)depth;
    h->call_idx = 0;
    use_lane(h, 0);
    return B200RX_OK;
}

int b200rx_join(b200rx_handle *h, uint32_t calls_back) { return b200rx_join_on(h, calls_back, h ? (void *)h->stream : nullptr); }

int b200rx_join_on(b200rx_handle *h, uint32_t calls_back, void *cuda_stream)
{
    if (!h) return B200RX_E_ARG;
    cudaStream_t target = (cudaStream_t)cuda_stream;
    if (h->depth <= 1) { // everything runs in order on the handle's stream: another stream waits for what is queued there
        if (target == h->stream) return B200RX_OK;
        CU(h, cudaSetDevice(h->device));
        if (!h->ev_in) CU(h, cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
        CU(h, cudaEventRecord(h->ev_in, h->stream));
        CU(h, cudaStreamWaitEvent(target, h->ev_in, 0));
        return B200RX_OK;
    }
    if (calls_back >= h->depth) return fail(h, B200RX_E_ARG, "b200rx_join: calls_back must be < pipeline depth");
    if (h->call_idx < (uint64_t)calls_back + 1) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    if (calls_back == 0) {
        // everything issued so far
        for (uint32_t i = 0; i < h->depth; i++)
            if (h->lanes[i].used) CU(h, cudaStreamWaitEvent(target, h->lanes[i].done, 0));
    } else {
        // exactly the call issued calls_back calls before the latest one (its lane has not been reused yet)
        const uint64_t c = h->call_idx - 1 - calls_back;
        CU(h, cudaStreamWaitEvent(target, h->lanes[c % h->depth].done, 0));
    }
    return B200RX_OK;
}

int b200rx_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return B200RX_E_ARG;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { *ptr = nullptr; return fail(nullptr, B200RX_E_NOMEM, "b200rx_host_alloc", e); }
    return B200RX_OK;
}

int b200rx_host_free(void *ptr)
{
    if (!ptr) return B200RX_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? B200RX_OK : B200RX_E_CUDA;
}

int b200rx_host_is_pinned(const void *ptr)
{
    if (!ptr) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return at.type == cudaMemoryTypeHost ? 1 : 0;
}

int b200rx_device_counters(b200rx_handle *h, void **dev_ptr)
{
    if (!h || !dev_ptr) return B200RX_E_ARG;
    *dev_ptr = h->counters;
    return B200RX_OK;
}

int b200rx_copy_counters(b200rx_handle *h, void *dst_dev)
{
    if (!h || !dst_dev) return B200RX_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (h->depth > 1 && h->call_idx > 0) {
        b200rx_handle::Lane &lane = h->lanes[(h->call_idx - 1) % h->depth];
        CU(h, cudaMemcpyAsync(dst_dev, lane.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, lane.stream));
        CU(h, cudaEventRecord(lane.done, lane.stream)); // the call is complete, for b200rx_join*, when its counters have been copied
        return B200RX_OK;
    }
    CU(h, cudaMemcpyAsync(dst_dev, h->counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
    return B200RX_OK;
}

uint64_t b200rx_launch_count(const b200rx_handle *h) { return h ? h->launches : 0; }
uint32_t b200rx_max_steps(const b200rx_handle *h) { return h ? h->max_steps : 0; }

size_t b200rx_sample_bytes(const b200rx_handle *h) { return h ? sample_bytes(h->fmt) : 0; }

namespace {

struct OutPtrs { uint8_t *payload; uint32_t stride; uint16_t *len; uint8_t *rate; uint8_t *status; };

// K1 -> K2 -> K3 for frames [off, off + n) of the batch on stream s; ev (4 events) optional.
int launch_range(b200rx_handle *h, const Tuning &tn, cudaStream_t s, uint32_t off, uint32_t n, const void *iq_dev, uint64_t iq_samples,
                 const uint64_t *lts1_dev, const uint32_t *avail_dev, const OutPtrs &o, const b200rx_debug *dbg,
                 cudaEvent_t *ev, const FrameRot *rot_dev = nullptr, const uint32_t *n_live_dev = nullptr,
                 const uint8_t *select_dev = nullptr)
{
    const size_t S = h->max_steps;
    FrontendArgs fa{};
    fa.iq = iq_dev;
    fa.fmt = h->fmt;
    fa.scale = h->scale;
    fa.iq_samples = iq_samples;
    fa.lts1 = lts1_dev + off;
    fa.avail = avail_dev + off;
    fa.n_frames = n;
    fa.desc = h->desc + off;
    fa.bm = h->bm + (size_t)off * S;
    fa.bm_stride = h->max_steps;
    fa.max_steps = h->max_steps;
    fa.max_len = h->limits.max_payload_bytes;
    fa.emit_pairs = tn.acs_gen == 3;
    fa.rot = rot_dev ? rot_dev + off : nullptr;
    fa.n_live = n_live_dev;
    fa.select = select_dev ? select_dev + off : nullptr;
    fa.sm_count = tn.sm_count;
    if (dbg) {
        fa.dbg_eq = dbg->equalized ? reinterpret_cast<double2 *>(dbg->equalized) + (size_t)off * dbg->eq_vectors * 48 : nullptr;
        fa.dbg_eq_vectors = dbg->eq_vectors;
        fa.dbg_depunct = dbg->depunct ? dbg->depunct + (size_t)off * dbg->depunct_stride : nullptr;
        fa.dbg_depunct_stride = dbg->depunct_stride;
    }
    if (ev) CU(h, cudaEventRecord(ev[0], s));
    if (tn.fe_split && tn.acs_gen == 3) { // header pass (descriptor + H^-1 per frame), then the data symbols
        fa.header_only = 2;
        fa.hinv_out = h->hinv + (size_t)off * 64;
        CU(h, launch_frontend(fa, s));
        CU(h, launch_frontend_data(fa, s));
        h->launches++;
    } else {
        CU(h, launch_frontend(fa, s));
    }
    if (ev) CU(h, cudaEventRecord(ev[1], s));
    if (tn.acs_gen == 3)
        CU(h, launch_viterbi_acs3(h->desc + off, reinterpret_cast<const uint8_t *>(h->bm + (size_t)off * S), 4 * S,
                                  h->dec + (size_t)off * 2 * S, 2 * h->max_steps, n, false, tn, s));
    else
        CU(h, launch_viterbi_acs(h->desc + off, h->bm + (size_t)off * S, h->max_steps, h->dec + (size_t)off * 2 * S,
                                 2 * h->max_steps, n, tn, s));
    if (ev) CU(h, cudaEventRecord(ev[2], s));
    TracebackArgs ta{};
    ta.desc = h->desc + off;
    ta.dec = h->dec + (size_t)off * 2 * S;
    ta.dec_stride = 2 * h->max_steps;
    ta.n_frames = n;
    ta.raw_mode = 0;
    ta.payload = o.payload ? o.payload + (size_t)off * o.stride : nullptr;
    ta.payload_stride = o.stride;
    ta.payload_len = o.len ? o.len + off : nullptr;
    ta.rate_out = o.rate ? o.rate + off : nullptr;
    ta.status_out = o.status + off;
    ta.counters = h->counters;
    if (dbg) {
        ta.dbg_decoded = dbg->decoded ? dbg->decoded + (size_t)off * dbg->decoded_stride : nullptr;
        ta.dbg_decoded_stride = dbg->decoded_stride;
        ta.dbg_field = dbg->header_field ? dbg->header_field + off : nullptr;
    }
    CU(h, launch_traceback(ta, tn, s));
    if (ev) CU(h, cudaEventRecord(ev[3], s));
    h->launches += 3;
    return B200RX_OK;
}

} // namespace

int b200rx_decode_batch_dev(b200rx_handle *h, const void *iq_dev, uint64_t iq_samples,
                            const uint64_t *lts1_index_dev, const uint32_t *avail_dev, uint32_t n_frames,
                            uint8_t *payload_out_dev, uint32_t payload_stride,
                            uint16_t *payload_len_dev, uint8_t *rate_out_dev, uint8_t *status_dev,
                            const b200rx_debug *dbg)
{
    if (!h) return B200RX_E_ARG;
    if (!iq_dev || !lts1_index_dev || !avail_dev || !status_dev)
        return fail(h, B200RX_E_ARG, "b200rx_decode_batch_dev: null argument");
    if (n_frames > h->limits.max_frames) return fail(h, B200RX_E_ARG, "b200rx_decode_batch_dev: n_frames exceeds max_frames");
    if (n_frames == 0) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    { int rcq = quiesce_host(h); if (rcq != B200RX_OK) return rcq; }
    cudaStream_t s = h->stream;
    b200rx_handle::Lane *lane = nullptr;
    if (h->depth > 1) {
        lane = &h->lanes[h->call_idx % h->depth];
        CU(h, cudaEventRecord(h->ev_in, h->stream)); // the caller's inputs are ready at this point of its stream
        CU(h, cudaStreamWaitEvent(lane->stream, h->ev_in, 0));
        s = lane->stream;
        use_lane(h, (int)(h->call_idx % h->depth));
    }

    CU(h, cudaMemsetAsync(h->counters, 0, 8 * sizeof(unsigned long long), s));
    cudaEvent_t *ev = call_events(h);
    const OutPtrs o{payload_out_dev, payload_stride, payload_len_dev, rate_out_dev, status_dev};
    int rc = launch_range(h, h->tn, s, 0, n_frames, iq_dev, iq_samples, lts1_index_dev, avail_dev, o, dbg, ev);
    if (rc != B200RX_OK) return rc;
    if (ev == h->ev) h->ev_valid = true;
    if (lane) {
        CU(h, cudaEventRecord(lane->done, s));
        lane->used = true;
    }
    h->call_idx++;
    return B200RX_OK;
}

namespace {

// scratch + small staging arrays of host slot k (slot 0 got them in b200rx_create)
int ensure_host_slot(b200rx_handle *h, int k)
{
    b200rx_handle::HostSlot &S = h->hs[k];
    if (!S.done) CU(h, cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
    if (S.desc && S.d_status) return B200RX_OK;
    const size_t nf = h->limits.max_frames;
    // all or nothing: a slot that could not get every buffer keeps none, so a later call retries instead of launching
    // kernels on null pointers
    ScratchPtrs sp;
    cudaError_t e = alloc_scratch(h, sp);
    uint64_t *d_lts1 = nullptr; uint32_t *d_avail = nullptr; uint16_t *d_len = nullptr; uint8_t *d_rate = nullptr, *d_status = nullptr;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void **)&d_lts1, nf * sizeof(uint64_t));
    A((void **)&d_avail, nf * sizeof(uint32_t));
    A((void **)&d_len, nf * sizeof(uint16_t));
    A((void **)&d_rate, nf);
    A((void **)&d_status, nf);
    if (e != cudaSuccess) {
        free_scratch(sp);
        cudaFree(d_lts1); cudaFree(d_avail); cudaFree(d_len); cudaFree(d_rate); cudaFree(d_status);
        (void)cudaGetLastError();
        return fail(h, B200RX_E_NOMEM, "b200rx_submit_batch: scratch for another call in flight", e);
    }
    S.desc = sp.desc; S.bm = sp.bm; S.dec = sp.dec; S.counters = sp.counters; S.hinv = sp.hinv;
    S.d_lts1 = d_lts1; S.d_avail = d_avail; S.d_len = d_len; S.d_rate = d_rate; S.d_status = d_status;
    return B200RX_OK;
}

int wait_slot(b200rx_handle *h, b200rx_handle::HostSlot &S)
{
    if (S.busy) {
        CU(h, cudaEventSynchronize(S.done));
        S.busy = false;
    }
    return B200RX_OK;
}

} // namespace

int b200rx_wait(b200rx_handle *h, uint64_t ticket)
{
    if (!h) return B200RX_E_ARG;
    CU(h, cudaSetDevice(h->device));
    for (auto &S : h->hs)
        if (S.busy && (ticket == 0 || S.ticket == ticket)) {
            int rc = wait_slot(h, S);
            if (rc != B200RX_OK) return rc;
        }
    return B200RX_OK;
}

int b200rx_decode_batch(b200rx_handle *h, const void *iq, uint64_t iq_samples,
                        const uint64_t *lts1_index, const uint32_t *avail, uint32_t n_frames,
                        uint8_t *payload_out, uint32_t payload_stride,
                        uint16_t *payload_len, uint8_t *rate_out, uint8_t *status)
{
    uint64_t ticket = 0;
    int rc = b200rx_submit_batch(h, iq, iq_samples, lts1_index, avail, n_frames, payload_out, payload_stride, payload_len,
                                 rate_out, status, &ticket);
    if (rc != B200RX_OK) return rc;
    return b200rx_wait(h, ticket);
}

int b200rx_submit_batch(b200rx_handle *h, const void *iq, uint64_t iq_samples,
                        const uint64_t *lts1_index, const uint32_t *avail, uint32_t n_frames,
                        uint8_t *payload_out, uint32_t payload_stride,
                        uint16_t *payload_len, uint8_t *rate_out, uint8_t *status, uint64_t *ticket)
try {
    if (!h) return B200RX_E_ARG;
    if (!iq || !lts1_index || !avail || !status || !ticket) return fail(h, B200RX_E_ARG, "b200rx_submit_batch: null argument");
    if (n_frames > h->limits.max_frames) return fail(h, B200RX_E_ARG, "b200rx_submit_batch: n_frames exceeds max_frames");
    *ticket = 0;
    if (n_frames == 0) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    if (h->depth > 1) { // the device-buffer lanes share scratch set 0 with host slot 0: drain them
        for (int i = 0; i < B200RX_MAX_PIPELINE_DEPTH; i++)
            if (h->lanes[i].stream && h->lanes[i].used) CU(h, cudaStreamSynchronize(h->lanes[i].stream));
    }
    // slot: round robin; a slot still in flight is waited for (at most B200RX_MAX_INFLIGHT calls overlap)
    const int k = (int)(h->next_ticket % B200RX_MAX_INFLIGHT);
    b200rx_handle::HostSlot &S = h->hs[k];
    { int rcw = wait_slot(h, S); if (rcw != B200RX_OK) return rcw; }
    { int rce = ensure_host_slot(h, k); if (rce != B200RX_OK) return rce; }
    cudaStream_t s = h->stream;
    h->desc = S.desc; h->bm = S.bm; h->dec = S.dec; h->counters = S.counters; h->hinv = S.hinv;

    const size_t bps = sample_bytes(h->fmt);
    const size_t iq_bytes = (size_t)iq_samples * bps;
    if (iq_bytes > S.d_iq_cap) {
        if (S.d_iq) { cudaFree(S.d_iq); S.d_iq = nullptr; S.d_iq_cap = 0; } // the slot is idle: nothing reads it
        cudaError_t e = cudaMalloc((void **)&S.d_iq, iq_bytes);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_submit_batch: sample staging", e);
        S.d_iq_cap = iq_bytes;
    }
    const size_t pl_bytes = payload_out ? (size_t)n_frames * payload_stride : 0;
    if (pl_bytes > S.d_payload_cap) {
        if (S.d_payload) { cudaFree(S.d_payload); S.d_payload = nullptr; S.d_payload_cap = 0; }
        cudaError_t e = cudaMalloc((void **)&S.d_payload, pl_bytes);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_submit_batch: payload staging", e);
        S.d_payload_cap = pl_bytes;
    }
    CU(h, cudaMemcpyAsync(S.d_lts1, lts1_index, n_frames * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(S.d_avail, avail, n_frames * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemsetAsync(S.counters, 0, 8 * sizeof(unsigned long long), s));
    if (pl_bytes) CU(h, cudaMemsetAsync(S.d_payload, 0, pl_bytes, s)); // bytes behind a frame's LENGTH reach the caller as zeros,
                                                                        // not as what an earlier call left in the staging
    const OutPtrs o{payload_out ? S.d_payload : nullptr, payload_stride, S.d_len, S.d_rate, S.d_status};
    auto finish = [&](cudaStream_t last) -> int {
        CU(h, cudaEventRecord(S.done, last));
        S.busy = true;
        S.ticket = h->next_ticket++;
        *ticket = S.ticket;
        return B200RX_OK;
    };

    // Chunked pipeline: the samples of chunk i+1 cross PCIe while chunk i is decoded (kernels of consecutive
    // chunks rotate over four streams so that their Viterbi kernels overlap) and chunk i-1's results go
    // back.  Needs the frames in stream order (lts1_index non-decreasing); otherwise one copy, one batch.
    // The copy is the bottleneck, so what matters for one call on its own is how long the decode of the LAST chunk
    // takes after its samples have landed: chunks start at CH frames and halve towards the end (down to CH_MIN).
    const uint32_t CH = (uint32_t)(h->tn.h2d_chunk >= 32 ? h->tn.h2d_chunk : 1024);
    const uint32_t CH_MIN = (uint32_t)(h->tn.h2d_chunk_min >= 16 ? h->tn.h2d_chunk_min : 256);
    // Pinned caller buffer: the GPU can pull the useful samples itself (ingest.cu) instead of a DMA copy of everything.
    // pull_mode 0: always DMA-copy, 1: pull whenever the buffer is pinned, 2: alternate, -1: by format (narrow formats:
    // the DMA engine wins)
    const int pull_mode = h->fmt == FMT_TAGGED ? 0 : (h->tn.pull_mode >= 0 ? h->tn.pull_mode : (h->fmt == FMT_FC64 ? 1 : 0));
    const bool pull_wanted = pull_mode != 0;
    const void *iq_mapped = nullptr;
    if (pull_wanted) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, iq) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            iq_mapped = at.devicePointer;
        else
            (void)cudaGetLastError();
    }
    bool ordered = n_frames > 2 * CH_MIN;
    for (uint32_t f = 1; ordered && f < n_frames; f++) ordered = lts1_index[f] >= lts1_index[f - 1];
    if (!ordered) {
        CU(h, cudaMemcpyAsync(S.d_iq, iq, iq_bytes, cudaMemcpyHostToDevice, s));
        int rc = launch_range(h, h->tn, s, 0, n_frames, S.d_iq, iq_samples, S.d_lts1, S.d_avail, o, nullptr, nullptr);
        if (rc != B200RX_OK) return rc;
        if (payload_out) CU(h, cudaMemcpyAsync(payload_out, S.d_payload, pl_bytes, cudaMemcpyDeviceToHost, s));
        if (payload_len) CU(h, cudaMemcpyAsync(payload_len, S.d_len, n_frames * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
        if (rate_out) CU(h, cudaMemcpyAsync(rate_out, S.d_rate, n_frames, cudaMemcpyDeviceToHost, s));
        CU(h, cudaMemcpyAsync(status, S.d_status, n_frames, cudaMemcpyDeviceToHost, s));
        return finish(s);
    }
    std::vector<uint32_t> bound(1, 0u); // chunk c = frames [bound[c], bound[c+1])
    for (uint32_t f = 0; f < n_frames;) {
        const uint32_t rest = n_frames - f;
        uint32_t take = CH;
        if (rest <= CH) take = rest <= CH_MIN ? rest : (rest / 2 > CH_MIN ? rest / 2 : CH_MIN);
        f += take;
        bound.push_back(f);
    }
    const uint32_t n_chunks = (uint32_t)bound.size() - 1;
    while (S.pipe_ev.size() < 2 * (size_t)n_chunks + 1) {
        cudaEvent_t e;
        CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        S.pipe_ev.push_back(e);
    }
    cudaEvent_t ev_start = S.pipe_ev[2 * n_chunks];
    CU(h, cudaEventRecord(ev_start, s)); // small arrays and counter reset queued
    CU(h, cudaStreamWaitEvent(h->copy_stream, ev_start, 0));
    CU(h, cudaStreamWaitEvent(h->pull_stream, ev_start, 0));
    for (int i = 0; i < 7; i++) CU(h, cudaStreamWaitEvent(h->aux_stream[i], ev_start, 0));
    uint64_t copied = 0; // samples [0, copied) are on their way
    // The link is the bottleneck here; what the kernels owe it is a short tail behind the last chunk's samples.  A chunk of
    // a few hundred frames is decoded fastest by the generation-2 ACS kernel with 16 lanes per frame (0.7 ms for 12 096
    // steps however few frames; generation 3 needs 1.4-1.7 ms for a lone chunk and wins only once the GPU is full).
    Tuning tn_chunk = h->tn;
    if (h->tn.acs_gen == 3 && CH <= 2048) tn_chunk.acs_gen = 2;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t f0 = bound[c], f1 = bound[c + 1];
        uint64_t hi = 0;
        for (uint32_t f = f0; f < f1; f++) {
            uint64_t e = lts1_index[f] + avail[f];
            if (e > iq_samples) e = iq_samples;
            if (e > hi) hi = e;
        }
        uint64_t lo = lts1_index[f0] < iq_samples ? lts1_index[f0] : iq_samples;
        if (lo < copied) lo = copied;
        const bool pull_this = iq_mapped && (pull_mode < 2 || (c % (uint32_t)pull_mode) != 0); // mode k >= 2: every k-th chunk goes
                                                                                        // through the DMA engine, the rest is pulled
        cudaStream_t in_stream = pull_this ? h->pull_stream : h->copy_stream;
        if (pull_this) {
            CU(h, launch_pull(iq_mapped, S.d_iq, h->fmt, iq_samples, S.d_lts1 + f0, S.d_avail + f0, f1 - f0, h->sm_count,
                              in_stream));
            h->launches++;
        } else if (hi > lo) {
            CU(h, cudaMemcpyAsync(S.d_iq + lo * bps, (const uint8_t *)iq + lo * bps, (size_t)(hi - lo) * bps,
                                  cudaMemcpyHostToDevice, in_stream));
            if (pull_mode < 2) copied = hi;
        }
        cudaEvent_t ev_in = S.pipe_ev[2 * c], ev_out = S.pipe_ev[2 * c + 1];
        CU(h, cudaEventRecord(ev_in, in_stream));
        cudaStream_t cs = (c & 7) ? h->aux_stream[(c & 7) - 1] : s;
        CU(h, cudaStreamWaitEvent(cs, ev_in, 0));
        int rc = launch_range(h, tn_chunk, cs, f0, f1 - f0, S.d_iq, iq_samples, S.d_lts1, S.d_avail, o, nullptr, nullptr);
        if (rc != B200RX_OK) return rc;
        CU(h, cudaEventRecord(ev_out, cs));
        CU(h, cudaStreamWaitEvent(h->d2h_stream, ev_out, 0));
        const uint32_t nf = f1 - f0;
        if (payload_out)
            CU(h, cudaMemcpyAsync(payload_out + (size_t)f0 * payload_stride, S.d_payload + (size_t)f0 * payload_stride,
                                  (size_t)nf * payload_stride, cudaMemcpyDeviceToHost, h->d2h_stream));
        if (payload_len) CU(h, cudaMemcpyAsync(payload_len + f0, S.d_len + f0, nf * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->d2h_stream));
        if (rate_out) CU(h, cudaMemcpyAsync(rate_out + f0, S.d_rate + f0, nf, cudaMemcpyDeviceToHost, h->d2h_stream));
        CU(h, cudaMemcpyAsync(status + f0, S.d_status + f0, nf, cudaMemcpyDeviceToHost, h->d2h_stream));
    }
    // every chunk's kernels are ordered before its copies on d2h_stream, so its tail is the end of the call
    return finish(h->d2h_stream);
}
catch (...) { return B200RX_E_NOMEM; } // std::bad_alloc must not cross the C boundary

namespace {

int ensure_sync_scratch(b200rx_handle *h, int lane)
{
    const size_t nf = h->limits.max_frames;
    h->sy_ev_cap = (uint32_t)(2 * nf + 64);
    cudaError_t e = cudaSuccess;
    if (!h->sy_summary_host) e = cudaHostAlloc((void **)&h->sy_summary_host, sizeof(SyncSummary), cudaHostAllocDefault);
    b200rx_handle::SyncScratch &y = h->sy[lane];
    if (e == cudaSuccess && !y.summary) {
        auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
        size_t ev_pow2 = 64; // the tagged scan sorts the list in place (bitonic): room for the next power of two
        while (ev_pow2 < h->sy_ev_cap) ev_pow2 *= 2;
        A((void **)&y.ev_x, ev_pow2 * sizeof(uint64_t));
        A((void **)&y.ev_count, 2 * sizeof(uint32_t));
        A((void **)&y.rec, h->sy_ev_cap * sizeof(SyncRec));
        A((void **)&y.lts1, nf * sizeof(uint64_t));
        A((void **)&y.avail, nf * sizeof(uint32_t));
        A((void **)&y.rot, nf * sizeof(FrameRot));
        A((void **)&y.phase, nf * sizeof(double));
        A((void **)&y.summary, sizeof(SyncSummary));
    }
    if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "sync scratch", e);
    return B200RX_OK;
}

// detector + timing sync on stream s into scratch set `lane` (asynchronous)
int launch_sync_lane(b200rx_handle *h, cudaStream_t s, int lane, const void *iq_dev, uint64_t n_samples, double phase_in,
                     uint8_t *tags_dev)
{
    if (h->fmt == FMT_TAGGED) return fail(h, B200RX_E_ARG, "tagged samples have been synchronised already: not a raw capture");
    int rc = ensure_sync_scratch(h, lane);
    if (rc != B200RX_OK) return rc;
    b200rx_handle::SyncScratch &y = h->sy[lane];
    const uint32_t n_ctas = sync_cta_count(n_samples);
    if (n_ctas > y.cta_cap) { // grow-only; work queued on this lane's stream may still read the old list
        if (y.cta_ev) {
            CU(h, cudaStreamSynchronize(s));
            cudaFree(y.cta_ev); cudaFree(y.cta_cnt);
            y.cta_ev = nullptr; y.cta_cnt = nullptr; y.cta_cap = 0;
        }
        cudaError_t e = cudaMalloc((void **)&y.cta_ev, (size_t)n_ctas * sizeof(CtaEvents));
        if (e == cudaSuccess) e = cudaMalloc((void **)&y.cta_cnt, n_ctas);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "sync scratch (detector event lists)", e);
        y.cta_cap = n_ctas;
    }
    SyncArgs a{};
    a.iq = iq_dev;
    a.fmt = h->fmt;
    a.scale = h->scale;
    a.n_samples = n_samples;
    a.rot_in = make_double2(cos(phase_in), sin(phase_in)); // timing_sync.cpp:124
    a.max_frames = h->limits.max_frames;
    a.tags = tags_dev;
    a.ev_x = y.ev_x; a.ev_count = y.ev_count; a.ev_cap = h->sy_ev_cap;
    a.rec = y.rec; a.cta_ev = y.cta_ev; a.cta_cnt = y.cta_cnt;
    if (!h->origins.empty() && h->origins.size() <= 16) { // consumed by this call; a short list rides in the kernel arguments
        a.n_inline = (uint32_t)h->origins.size();
        for (uint32_t o = 0; o < a.n_inline; o++) a.origins_inline[o] = h->origins[o];
        h->origins.clear();
    }
    if (!h->origins.empty()) {
        const uint32_t no = (uint32_t)h->origins.size();
        if (no > h->d_origins_cap[lane]) {
            if (h->d_origins[lane]) { CU(h, cudaStreamSynchronize(s)); cudaFree(h->d_origins[lane]); h->d_origins[lane] = nullptr; }
            cudaError_t e = cudaMalloc((void **)&h->d_origins[lane], (size_t)no * 2 * sizeof(int64_t));
            if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "sync scratch (buffer origins)", e);
            h->d_origins_cap[lane] = no * 2;
        }
        // pageable source: the copy is staged before the call returns, the vector may be reused right away
        CU(h, cudaMemcpyAsync(h->d_origins[lane], h->origins.data(), no * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        a.origins = h->d_origins[lane];
        a.n_origins = no;
        h->origins.clear();
    }
    a.lts1 = y.lts1; a.avail = y.avail; a.rot = y.rot; a.phase = y.phase;
    a.summary = y.summary;
    CU(h, launch_sync(a, s));
    h->launches += n_samples ? 4 : 1; // detect, scan, lts_sync, build_frames
    return B200RX_OK;
}

// summary of scratch set `lane` to the host; returns when everything queued on s so far has finished
int fetch_summary(b200rx_handle *h, cudaStream_t s, int lane, double phase_in, b200rx_sync_result *res)
{
    CU(h, cudaMemcpyAsync(h->sy_summary_host, h->sy[lane].summary, sizeof(SyncSummary), cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    *res = *h->sy_summary_host;
    if (!res->phase_valid) res->last_phase = phase_in;
    return B200RX_OK;
}

int drain_lanes(b200rx_handle *h)
{
    { int rcq = quiesce_host(h); if (rcq != B200RX_OK) return rcq; }
    if (h->depth > 1) { // not pipelined: drain the lanes, work on scratch set 0
        int rcq = b200rx_synchronize(h);
        if (rcq != B200RX_OK) return rcq;
        use_lane(h, 0);
    }
    return B200RX_OK;
}

} // namespace

int b200rx_sync_dev(b200rx_handle *h, const void *iq_dev, uint64_t n_samples, double phase_in, uint8_t *tags_dev,
                    uint64_t *lts1_index_dev, uint32_t *avail_dev, double *phase_dev, b200rx_sync_result *res)
{
    if (!h) return B200RX_E_ARG;
    if ((!iq_dev && n_samples) || !res) return fail(h, B200RX_E_ARG, "b200rx_sync_dev: null argument");
    CU(h, cudaSetDevice(h->device));
    int rc = drain_lanes(h);
    if (rc != B200RX_OK) return rc;
    cudaStream_t s = h->stream;
    rc = launch_sync_lane(h, s, 0, iq_dev, n_samples, phase_in, tags_dev);
    if (rc != B200RX_OK) return rc;
    rc = fetch_summary(h, s, 0, phase_in, res);
    if (rc != B200RX_OK) return rc;
    const size_t nf = res->n_frames;
    if (nf) {
        const b200rx_handle::SyncScratch &y = h->sy[0];
        if (lts1_index_dev) CU(h, cudaMemcpyAsync(lts1_index_dev, y.lts1, nf * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        if (avail_dev) CU(h, cudaMemcpyAsync(avail_dev, y.avail, nf * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        if (phase_dev) CU(h, cudaMemcpyAsync(phase_dev, y.phase, nf * sizeof(double), cudaMemcpyDeviceToDevice, s));
        CU(h, cudaStreamSynchronize(s));
    }
    return B200RX_OK;
}

int b200rx_receive_dev(b200rx_handle *h, const void *iq_dev, uint64_t n_samples, double phase_in,
                       uint8_t *payload_out_dev, uint32_t payload_stride, uint16_t *payload_len_dev, uint8_t *rate_out_dev,
                       uint8_t *status_dev, uint64_t *lts1_out_dev, uint32_t *n_frames_dev, b200rx_sync_result *res)
{
    if (!h) return B200RX_E_ARG;
    if ((!iq_dev && n_samples) || !status_dev) return fail(h, B200RX_E_ARG, "b200rx_receive_dev: null argument");
    CU(h, cudaSetDevice(h->device));
    { int rcq = quiesce_host(h); if (rcq != B200RX_OK) return rcq; }
    cudaStream_t s = h->stream;
    b200rx_handle::Lane *lane = nullptr;
    int li = 0;
    if (h->depth > 1) {
        li = (int)(h->call_idx % h->depth);
        lane = &h->lanes[li];
        CU(h, cudaEventRecord(h->ev_in, h->stream)); // the caller's samples are ready at this point of its stream
        CU(h, cudaStreamWaitEvent(lane->stream, h->ev_in, 0));
        s = lane->stream;
        use_lane(h, li);
    }
    const uint32_t mf = h->limits.max_frames;
    CU(h, cudaMemsetAsync(h->counters, 0, 8 * sizeof(unsigned long long), s));
    int rc = launch_sync_lane(h, s, li, iq_dev, n_samples, phase_in, nullptr);
    if (rc != B200RX_OK) return rc;
    const b200rx_handle::SyncScratch &y = h->sy[li];
    cudaEvent_t *ev = call_events(h);
    const OutPtrs o{payload_out_dev, payload_stride, payload_len_dev, rate_out_dev, status_dev};
    rc = launch_range(h, h->tn, s, 0, mf, iq_dev, n_samples, y.lts1, y.avail, o, nullptr, ev, y.rot, &y.summary->n_frames);
    if (rc != B200RX_OK) return rc;
    if (ev == h->ev) h->ev_valid = true;
    if (lts1_out_dev) CU(h, cudaMemcpyAsync(lts1_out_dev, y.lts1, (size_t)mf * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    if (n_frames_dev) CU(h, cudaMemcpyAsync(n_frames_dev, &y.summary->n_frames, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    if (lane) {
        CU(h, cudaEventRecord(lane->done, s));
        lane->used = true;
    }
    h->call_idx++;
    if (res) return fetch_summary(h, s, li, phase_in, res);
    return B200RX_OK;
}

int b200rx_receive(b200rx_handle *h, const void *iq, uint64_t n_samples, double phase_in, uint8_t *payload_out,
                   uint32_t payload_stride, uint16_t *payload_len, uint8_t *rate_out, uint8_t *status, uint64_t *lts1_out,
                   b200rx_sync_result *res)
{
    if (!h) return B200RX_E_ARG;
    if ((!iq && n_samples) || !status || !res) return fail(h, B200RX_E_ARG, "b200rx_receive: null argument");
    CU(h, cudaSetDevice(h->device));
    int rc = drain_lanes(h);
    if (rc != B200RX_OK) return rc;
    cudaStream_t s = h->stream;
    const size_t iq_bytes = (size_t)n_samples * sample_bytes(h->fmt);
    if (iq_bytes > h->hs[0].d_iq_cap) {
        if (h->hs[0].d_iq) { CU(h, cudaStreamSynchronize(s)); cudaFree(h->hs[0].d_iq); h->hs[0].d_iq = nullptr; h->hs[0].d_iq_cap = 0; }
        cudaError_t e = cudaMalloc((void **)&h->hs[0].d_iq, iq_bytes);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_receive: sample staging", e);
        h->hs[0].d_iq_cap = iq_bytes;
    }
    const size_t pl_cap = payload_out ? (size_t)h->limits.max_frames * payload_stride : 0;
    if (pl_cap > h->hs[0].d_payload_cap) {
        if (h->hs[0].d_payload) { CU(h, cudaStreamSynchronize(s)); cudaFree(h->hs[0].d_payload); h->hs[0].d_payload = nullptr; h->hs[0].d_payload_cap = 0; }
        cudaError_t e = cudaMalloc((void **)&h->hs[0].d_payload, pl_cap);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_receive: payload staging", e);
        h->hs[0].d_payload_cap = pl_cap;
    }
    if (iq_bytes) CU(h, cudaMemcpyAsync(h->hs[0].d_iq, iq, iq_bytes, cudaMemcpyHostToDevice, s));
    if (pl_cap) CU(h, cudaMemsetAsync(h->hs[0].d_payload, 0, pl_cap, s));
    const int li = h->depth > 1 ? (int)(h->call_idx % h->depth) : 0; // scratch set the next call uses
    rc = b200rx_receive_dev(h, h->hs[0].d_iq, n_samples, phase_in, payload_out ? h->hs[0].d_payload : nullptr, payload_stride, h->hs[0].d_len,
                            h->hs[0].d_rate, h->hs[0].d_status, nullptr, nullptr, res);
    if (rc != B200RX_OK) return rc;
    const size_t nf = res->n_frames;
    if (nf) {
        if (payload_out) CU(h, cudaMemcpyAsync(payload_out, h->hs[0].d_payload, nf * payload_stride, cudaMemcpyDeviceToHost, s));
        if (payload_len) CU(h, cudaMemcpyAsync(payload_len, h->hs[0].d_len, nf * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
        if (rate_out) CU(h, cudaMemcpyAsync(rate_out, h->hs[0].d_rate, nf, cudaMemcpyDeviceToHost, s));
        CU(h, cudaMemcpyAsync(status, h->hs[0].d_status, nf, cudaMemcpyDeviceToHost, s));
        if (lts1_out) CU(h, cudaMemcpyAsync(lts1_out, h->sy[li].lts1, nf * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    }
    CU(h, cudaStreamSynchronize(s));
    return B200RX_OK;
}

// ---- two-phase passes (include/b200rx.h: b200rx_pass_*) ----
namespace {

// frame list of a scanned capture, written straight into host-mapped memory (one posted PCIe write per frame instead of
// four device->host copies in front of the synchronisation the caller waits on)
__global__ void pack_pass_kernel(const SyncSummary *summary, const uint64_t *lts1, const uint32_t *avail, const FrameDesc *desc,
                                 const FrameRot *rot, const double *phase, uint32_t cap, SyncSummary *out_summary,
                                 b200rx_pass_frame *out)
{
    const uint32_t n = summary->n_frames < cap ? summary->n_frames : cap;
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f == 0) *out_summary = *summary;
    if (f >= n) return;
    const FrameDesc d = desc[f];
    b200rx_pass_frame o;
    o.lts1 = lts1[f];
    o.avail = avail[f];
    o.length = d.length;
    o.rate = d.rate;
    o.status = d.status;
    o.sts_end = rot ? rot[f].from : 0ull;
    o.phase = phase ? phase[f] : 0.0;
    out[f] = o;
}

// ---- tagged streams (b200rx_pass_scan_tagged): find the LTS1 tags, list the frames ----
// Every sample's tag is looked at once; the few positions found are appended in any order ...
__global__ void tag_find_kernel(const char *stream, uint64_t n_samples, uint64_t *pos, uint32_t *count, uint32_t cap)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    const int tag = *reinterpret_cast<const int *>(stream + i * TAGGED_SAMPLE_BYTES + 16);
    if (tag != TAG_LTS1) return;
    const uint32_t k = atomicAdd(count, 1u);
    if (k < cap) pos[k] = i;
}

// ... and one CTA puts them in stream order (bitonic sort, in place) and writes the frame list: a frame reaches to the
// next LTS1 or to the end of the stream (fft_symbols.cpp:42-51).
__global__ void __launch_bounds__(1024) tag_frames_kernel(uint64_t *pos, const uint32_t *count, uint32_t cap, uint32_t max_frames,
                                                          uint64_t n_samples, uint64_t *lts1, uint32_t *avail, SyncSummary *summary)
{
    const uint32_t found = *count;
    const uint32_t n = found < cap ? found : cap;
    uint32_t n2 = 1;
    while (n2 < n) n2 *= 2;
    for (uint32_t i = n + threadIdx.x; i < n2; i += blockDim.x) pos[i] = ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k *= 2)
        for (uint32_t j = k / 2; j > 0; j /= 2) {
            for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const uint64_t a = pos[i], b = pos[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { pos[i] = b; pos[l] = a; }
                }
            }
            __syncthreads();
        }
    const uint32_t nf = n < max_frames ? n : max_frames;
    for (uint32_t f = threadIdx.x; f < nf; f += blockDim.x) {
        const uint64_t a = pos[f];
        const uint64_t e = (f + 1 < n) ? pos[f + 1] : n_samples;
        lts1[f] = a;
        avail[f] = (uint32_t)(e - a < 0xFFFFFFFFull ? e - a : 0xFFFFFFFFull);
    }
    if (threadIdx.x == 0) {
        SyncSummary s{};
        s.n_events = found;
        s.n_frames = nf;
        s.overflow = found - nf;
        *summary = s;
    }
}

int ensure_pass_lane(b200rx_handle *h, int li)
{
    b200rx_handle::PassLane &P = h->pl[li];
    if (P.h_select) return B200RX_OK;
    const size_t nf = h->limits.max_frames;
    cudaError_t e = cudaSuccess;
    if (!P.stream) e = cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess && !P.done) e = cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming);
    uint8_t *d_status = nullptr, *d_select = nullptr, *h_list = nullptr, *h_select = nullptr;
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_status, nf);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_select, nf);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&h_list, sizeof(SyncSummary) + nf * sizeof(b200rx_pass_frame), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&h_select, nf, cudaHostAllocDefault);
    void *dev = nullptr;
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&dev, h_list, 0);
    if (e != cudaSuccess) {
        cudaFree(d_status); cudaFree(d_select);
        if (h_list) cudaFreeHost(h_list);
        if (h_select) cudaFreeHost(h_select);
        (void)cudaGetLastError();
        return fail(h, B200RX_E_NOMEM, "b200rx_pass_open: lane staging", e);
    }
    P.d_status = d_status; P.d_select = d_select; P.h_list = h_list; P.h_list_dev = (uint8_t *)dev; P.h_select = h_select;
    return B200RX_OK;
}

} // namespace

int b200rx_pass_open(b200rx_handle *h)
{
    if (!h) return B200RX_E_ARG;
    CU(h, cudaSetDevice(h->device));
    { int rcq = quiesce_host(h); if (rcq != B200RX_OK) return rcq; }
    const int li = (int)(h->pass_count % h->depth);
    if (h->lanes[li].stream && h->lanes[li].used) { // a device-buffer call still on this lane's scratch
        CU(h, cudaStreamSynchronize(h->lanes[li].stream));
        h->lanes[li].used = false;
    }
    if (li == 0) CU(h, cudaStreamSynchronize(h->stream)); // scratch set 0 also serves the synchronous entry points
    int rc = ensure_pass_lane(h, li);
    if (rc != B200RX_OK) return rc;
    b200rx_handle::PassLane &P = h->pl[li];
    if (P.busy) {
        CU(h, cudaEventSynchronize(P.done));
        P.busy = false;
    }
    P.fill = 0;
    P.n_frames = 0;
    P.scanned = false;
    P.tagged = false;
    h->pass_lane = li;
    h->pass_count++;
    return B200RX_OK;
}

int b200rx_pass_put(b200rx_handle *h, const void *iq, uint64_t n_samples)
{
    if (!h) return B200RX_E_ARG;
    if (h->pass_lane < 0) return fail(h, B200RX_E_ARG, "b200rx_pass_put: no pass open");
    if (n_samples == 0) return B200RX_OK;
    if (!iq) return fail(h, B200RX_E_ARG, "b200rx_pass_put: null argument");
    b200rx_handle::PassLane &P = h->pl[h->pass_lane];
    if (P.scanned) return fail(h, B200RX_E_ARG, "b200rx_pass_put: the pass has been scanned already");
    CU(h, cudaSetDevice(h->device));
    const size_t bps = sample_bytes(h->fmt);
    const size_t need = (size_t)(P.fill + n_samples) * bps;
    if (need > P.d_iq_cap) { // grow (the lane is idle apart from this pass's own copies)
        size_t cap = P.d_iq_cap ? P.d_iq_cap : ((size_t)1 << 20);
        while (cap < need) cap *= 2;
        uint8_t *p = nullptr;
        cudaError_t e = cudaMalloc((void **)&p, cap);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_pass_put: sample staging", e);
        if (P.fill) CU(h, cudaMemcpyAsync(p, P.d_iq, (size_t)P.fill * bps, cudaMemcpyDeviceToDevice, P.stream));
        CU(h, cudaStreamSynchronize(P.stream));
        cudaFree(P.d_iq);
        P.d_iq = p;
        P.d_iq_cap = cap;
    }
    CU(h, cudaMemcpyAsync(P.d_iq + (size_t)P.fill * bps, iq, (size_t)n_samples * bps, cudaMemcpyHostToDevice, P.stream));
    P.fill += n_samples;
    return B200RX_OK;
}

namespace {

// Captures the launches of one scan (parameter copy, detector, event scan, LTS synchronisation, frame list, SIGNAL decode,
// frame-list export) on the lane's stream into a graph whose kernels take this call's scalars from the lane's ScanParams
// block.  A streaming caller scans every 4096 samples: replaying one graph costs one driver call instead of eight.
int build_scan_graph(b200rx_handle *h, int li)
{
    b200rx_handle::PassLane &P = h->pl[li];
    const size_t bps = sample_bytes(h->fmt);
    const uint64_t cap = P.d_iq_cap / bps;
    if (P.scan_exec) { cudaGraphExecDestroy(P.scan_exec); P.scan_exec = nullptr; }
    if (cap == 0) return B200RX_E_ARG;
    int rc = ensure_sync_scratch(h, li);
    if (rc != B200RX_OK) return rc;
    cudaStream_t s = P.stream;
    b200rx_handle::SyncScratch &y = h->sy[li];
    const uint32_t n_ctas = sync_cta_count(cap);
    if (n_ctas > y.cta_cap) {
        CU(h, cudaStreamSynchronize(s));
        cudaFree(y.cta_ev); cudaFree(y.cta_cnt);
        y.cta_ev = nullptr; y.cta_cnt = nullptr; y.cta_cap = 0;
        cudaError_t e = cudaMalloc((void **)&y.cta_ev, (size_t)n_ctas * sizeof(CtaEvents));
        if (e == cudaSuccess) e = cudaMalloc((void **)&y.cta_cnt, n_ctas);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "sync scratch (detector event lists)", e);
        y.cta_cap = n_ctas;
    }
    if (!P.h_sp) {
        cudaError_t e = cudaHostAlloc((void **)&P.h_sp, sizeof(ScanParams), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc((void **)&P.d_sp, sizeof(ScanParams));
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "scan parameter block", e);
        memset(P.h_sp, 0, sizeof(ScanParams));
    }
    use_lane(h, li);
    const uint32_t mf = h->limits.max_frames;
    SyncArgs a{};
    a.iq = P.d_iq; a.fmt = h->fmt; a.scale = h->scale; a.n_samples = cap; a.grid_samples = cap; a.sp = P.d_sp;
    a.rot_in = make_double2(1.0, 0.0);
    a.max_frames = mf; a.tags = nullptr;
    a.ev_x = y.ev_x; a.ev_count = y.ev_count; a.ev_cap = h->sy_ev_cap;
    a.rec = y.rec; a.cta_ev = y.cta_ev; a.cta_cnt = y.cta_cnt;
    a.lts1 = y.lts1; a.avail = y.avail; a.rot = y.rot; a.phase = y.phase; a.summary = y.summary;
    FrontendArgs fa{};
    fa.iq = P.d_iq; fa.fmt = h->fmt; fa.scale = h->scale; fa.iq_samples = cap; fa.sp = P.d_sp;
    fa.lts1 = y.lts1; fa.avail = y.avail;
    fa.n_frames = (uint32_t)((cap / 17 + 2 < (uint64_t)mf) ? cap / 17 + 2 : mf);
    fa.desc = h->desc; fa.bm = h->bm; fa.bm_stride = h->max_steps; fa.max_steps = h->max_steps;
    fa.max_len = h->limits.max_payload_bytes; fa.header_only = 2; fa.hinv_out = h->hinv; fa.rot = y.rot;
    fa.n_live = &y.summary->n_frames;
    SyncSummary *out_summary = reinterpret_cast<SyncSummary *>(P.h_list_dev);
    b200rx_pass_frame *out_frames = reinterpret_cast<b200rx_pass_frame *>(P.h_list_dev + sizeof(SyncSummary));

    CU(h, cudaStreamSynchronize(s));
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return fail(h, B200RX_E_CUDA, "scan graph: begin capture", e);
    e = cudaMemcpyAsync(P.d_sp, P.h_sp, sizeof(ScanParams), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_sync(a, s);
    if (e == cudaSuccess) e = launch_frontend(fa, s);
    if (e == cudaSuccess) {
        pack_pass_kernel<<<(mf + 127) / 128, 128, 0, s>>>(y.summary, y.lts1, y.avail, h->desc, y.rot, y.phase, mf, out_summary, out_frames);
        e = cudaGetLastError();
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(s, &graph);
    if (e == cudaSuccess) e = e_end;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&P.scan_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        P.scan_exec = nullptr;
        (void)cudaGetLastError();
        return fail(h, B200RX_E_CUDA, "scan graph: capture", e);
    }
    P.scan_iq = P.d_iq; P.scan_cap = cap; P.scan_fmt = h->fmt; P.scan_scale = h->scale;
    return B200RX_OK;
}

} // namespace

static int pass_scan_impl(b200rx_handle *h, bool tagged, double phase_in, b200rx_pass_frame *frames, uint32_t frames_cap,
                          b200rx_sync_result *res)
{
    if (!h) return B200RX_E_ARG;
    if (!res || (!frames && frames_cap)) return fail(h, B200RX_E_ARG, "b200rx_pass_scan: null argument");
    if (tagged != (h->fmt == FMT_TAGGED))
        return fail(h, B200RX_E_ARG, tagged ? "b200rx_pass_scan_tagged: the sample format must be B200RX_FMT_TAGGED_FC64"
                                            : "b200rx_pass_scan: tagged samples have been synchronised already (b200rx_pass_scan_tagged)");
    if (h->pass_lane < 0) return fail(h, B200RX_E_ARG, "b200rx_pass_scan: no pass open");
    const int li = h->pass_lane;
    b200rx_handle::PassLane &P = h->pl[li];
    if (P.scanned) return fail(h, B200RX_E_ARG, "b200rx_pass_scan: the pass has been scanned already");
    CU(h, cudaSetDevice(h->device));
    memset(res, 0, sizeof(*res));
    res->last_phase = phase_in;
    P.scanned = true;
    P.tagged = tagged;
    if (P.fill == 0) {
        h->origins.clear();
        return B200RX_OK;
    }
    cudaStream_t s = P.stream;
    use_lane(h, li);
    const uint32_t mf = h->limits.max_frames;
    int rc;
    if (tagged) {
        rc = ensure_sync_scratch(h, li);
        if (rc != B200RX_OK) return rc;
        const b200rx_handle::SyncScratch &t = h->sy[li];
        CU(h, cudaMemsetAsync(t.ev_count, 0, 2 * sizeof(uint32_t), s));
        tag_find_kernel<<<(unsigned)((P.fill + 255) / 256), 256, 0, s>>>(reinterpret_cast<const char *>(P.d_iq), P.fill, t.ev_x,
                                                                        t.ev_count, h->sy_ev_cap);
        tag_frames_kernel<<<1, 1024, 0, s>>>(t.ev_x, t.ev_count, h->sy_ev_cap, mf, P.fill, t.lts1, t.avail, t.summary);
        CU(h, cudaGetLastError());
        h->launches += 2;
    } else if (h->tn.scan_graph && h->origins.size() <= 16) {
        // the whole scan as one graph launch (see build_scan_graph)
        if (!P.scan_exec || P.scan_iq != P.d_iq || P.scan_cap != P.d_iq_cap / sample_bytes(h->fmt) || P.scan_fmt != h->fmt ||
            P.scan_scale != h->scale) {
            rc = build_scan_graph(h, li);
            if (rc != B200RX_OK) return rc;
        }
        ScanParams &sp = *P.h_sp;
        sp.n_samples = P.fill;
        sp.x_limit = P.fill > 160 ? P.fill - 160 : 0; // timing_sync.cpp:68: x < input.size() - CARRYOVER_LENGTH
        sp.rot_in = make_double2(cos(phase_in), sin(phase_in)); // timing_sync.cpp:124
        sp.n_origins = (uint32_t)h->origins.size();
        for (uint32_t o = 0; o < sp.n_origins; o++) sp.origins[o] = h->origins[o];
        h->origins.clear();
        CU(h, cudaGraphLaunch(P.scan_exec, s));
        h->launches += 7;
        CU(h, cudaStreamSynchronize(s));
        *res = *reinterpret_cast<const SyncSummary *>(P.h_list);
        if (!res->phase_valid) res->last_phase = phase_in;
        P.n_frames = res->n_frames < mf ? res->n_frames : mf;
        const uint32_t n_copy = P.n_frames < frames_cap ? P.n_frames : frames_cap;
        if (n_copy) memcpy(frames, P.h_list + sizeof(SyncSummary), (size_t)n_copy * sizeof(b200rx_pass_frame));
        return B200RX_OK;
    } else {
        rc = launch_sync_lane(h, s, li, P.d_iq, P.fill, phase_in, nullptr);
        if (rc != B200RX_OK) return rc;
    }
    const b200rx_handle::SyncScratch &y = h->sy[li];
    FrontendArgs fa{}; // SIGNAL decode of every frame found: descriptor with the full status logic
    fa.iq = P.d_iq; fa.fmt = h->fmt; fa.scale = h->scale; fa.iq_samples = P.fill;
    // frames start at distinct STS_END / LTS1 tags, at least 17 samples apart: the grid need not cover more slots than that
    const uint32_t slots = (uint32_t)((P.fill / 17 + 2 < (uint64_t)mf) ? P.fill / 17 + 2 : mf);
    fa.lts1 = y.lts1; fa.avail = y.avail; fa.n_frames = tagged ? mf : slots; fa.desc = h->desc; fa.bm = h->bm;
    fa.bm_stride = h->max_steps; fa.max_steps = h->max_steps; fa.max_len = h->limits.max_payload_bytes;
    fa.header_only = 2; fa.hinv_out = h->hinv; fa.rot = tagged ? nullptr : y.rot; fa.n_live = &y.summary->n_frames;
    CU(h, launch_frontend(fa, s));
    SyncSummary *out_summary = reinterpret_cast<SyncSummary *>(P.h_list_dev);
    b200rx_pass_frame *out_frames = reinterpret_cast<b200rx_pass_frame *>(P.h_list_dev + sizeof(SyncSummary));
    pack_pass_kernel<<<(mf + 127) / 128, 128, 0, s>>>(y.summary, y.lts1, y.avail, h->desc, tagged ? nullptr : y.rot,
                                                      tagged ? nullptr : y.phase, mf, out_summary, out_frames);
    CU(h, cudaGetLastError());
    h->launches += 2;
    CU(h, cudaStreamSynchronize(s));
    *res = *reinterpret_cast<const SyncSummary *>(P.h_list);
    if (!res->phase_valid) res->last_phase = phase_in;
    P.n_frames = res->n_frames < mf ? res->n_frames : mf;
    const uint32_t n_copy = P.n_frames < frames_cap ? P.n_frames : frames_cap;
    if (n_copy) memcpy(frames, P.h_list + sizeof(SyncSummary), (size_t)n_copy * sizeof(b200rx_pass_frame));
    return B200RX_OK;
}

int b200rx_pass_scan(b200rx_handle *h, double phase_in, b200rx_pass_frame *frames, uint32_t frames_cap, b200rx_sync_result *res)
{
    return pass_scan_impl(h, false, phase_in, frames, frames_cap, res);
}

int b200rx_pass_scan_tagged(b200rx_handle *h, b200rx_pass_frame *frames, uint32_t frames_cap, b200rx_sync_result *res)
{
    return pass_scan_impl(h, true, 0.0, frames, frames_cap, res);
}

int b200rx_pass_decode(b200rx_handle *h, const uint8_t *select, uint8_t *payload_out, uint32_t payload_stride, uint8_t *status,
                       uint64_t *ticket)
{
    if (!h) return B200RX_E_ARG;
    if (!select || !status || !ticket) return fail(h, B200RX_E_ARG, "b200rx_pass_decode: null argument");
    if (h->pass_lane < 0 || !h->pl[h->pass_lane].scanned) return fail(h, B200RX_E_ARG, "b200rx_pass_decode: no scanned pass");
    const int li = h->pass_lane;
    b200rx_handle::PassLane &P = h->pl[li];
    if (P.busy) return fail(h, B200RX_E_ARG, "b200rx_pass_decode: the pass is being decoded already");
    *ticket = 0;
    const uint32_t n = P.n_frames;
    uint32_t lo = n, hi = 0, n_sel = 0;
    for (uint32_t f = 0; f < n; f++)
        if (select[f]) {
            if (f < lo) lo = f;
            hi = f + 1;
            n_sel++;
        }
    if (n_sel == 0) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    cudaStream_t s = P.stream;
    use_lane(h, li);
    const size_t pl_bytes = payload_out ? (size_t)n * payload_stride : 0;
    if (pl_bytes > P.d_payload_cap) {
        cudaFree(P.d_payload); // idle: the lane's previous pass was waited for in b200rx_pass_open
        P.d_payload = nullptr; P.d_payload_cap = 0;
        cudaError_t e = cudaMalloc((void **)&P.d_payload, pl_bytes);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_pass_decode: payload staging", e);
        P.d_payload_cap = pl_bytes;
    }
    memcpy(P.h_select, select, n);
    CU(h, cudaMemcpyAsync(P.d_select, P.h_select, n, cudaMemcpyHostToDevice, s));
    if (payload_out) CU(h, cudaMemsetAsync(P.d_payload + (size_t)lo * payload_stride, 0, (size_t)(hi - lo) * payload_stride, s));
    CU(h, cudaMemsetAsync(h->counters, 0, 8 * sizeof(unsigned long long), s));
    // A pass holds a handful of frames unless the caller hands over a long capture; what it owes the caller is a short
    // tail behind the scan: the generation-2 ACS kernel with 16 lanes per frame takes 0.7 ms for 12 096 dependent steps
    // however few frames there are, generation 3 is built for a full GPU
    Tuning tn = h->tn;
    if (tn.acs_gen == 3 && n_sel <= 2048) tn.acs_gen = 2;
    if (tn.acs_gen == 2 && tn.acs_lb == 0 && n_sel <= 64) tn.acs_lb = 5; // a warp per frame: 0.65 instead of 0.68 ms for 12 096 steps
    const b200rx_handle::SyncScratch &y = h->sy[li];
    const OutPtrs o{payload_out ? P.d_payload : nullptr, payload_stride, nullptr, nullptr, P.d_status};
    int rc = launch_range(h, tn, s, lo, hi - lo, P.d_iq, P.fill, y.lts1, y.avail, o, nullptr, nullptr, P.tagged ? nullptr : y.rot,
                          nullptr, P.d_select);
    if (rc != B200RX_OK) return rc;
    if (payload_out)
        CU(h, cudaMemcpyAsync(payload_out + (size_t)lo * payload_stride, P.d_payload + (size_t)lo * payload_stride,
                              (size_t)(hi - lo) * payload_stride, cudaMemcpyDeviceToHost, s));
    CU(h, cudaMemcpyAsync(status + lo, P.d_status + lo, hi - lo, cudaMemcpyDeviceToHost, s));
    CU(h, cudaEventRecord(P.done, s));
    P.busy = true;
    P.ticket = h->pass_next_ticket++;
    *ticket = P.ticket;
    return B200RX_OK;
}

int b200rx_pass_poll(b200rx_handle *h, uint64_t ticket)
{
    if (!h) return B200RX_E_ARG;
    if (ticket == 0) return 1;
    for (auto &P : h->pl)
        if (P.busy && P.ticket == ticket) {
            cudaError_t e = cudaEventQuery(P.done);
            if (e == cudaErrorNotReady) return 0;
            if (e != cudaSuccess) return fail(h, B200RX_E_CUDA, "b200rx_pass_poll", e);
            P.busy = false;
            return 1;
        }
    return 1; // not in flight any more
}

int b200rx_pass_wait(b200rx_handle *h, uint64_t ticket)
{
    if (!h) return B200RX_E_ARG;
    for (auto &P : h->pl)
        if (P.busy && (ticket == 0 || P.ticket == ticket)) {
            CU(h, cudaEventSynchronize(P.done));
            P.busy = false;
        }
    return B200RX_OK;
}

int b200rx_decode_headers(b200rx_handle *h, const void *iq, uint64_t iq_samples, const uint64_t *lts1_index,
                          const uint32_t *avail, uint32_t n_frames, uint16_t *payload_len, uint8_t *rate_out,
                          uint8_t *status)
{
    if (!h) return B200RX_E_ARG;
    if (!iq || !lts1_index || !avail || !status) return fail(h, B200RX_E_ARG, "b200rx_decode_headers: null argument");
    if (n_frames > h->limits.max_frames) return fail(h, B200RX_E_ARG, "b200rx_decode_headers: n_frames exceeds max_frames");
    if (n_frames == 0) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    { int rcq = drain_lanes(h); if (rcq != B200RX_OK) return rcq; } // not pipelined: work on scratch set 0
    cudaStream_t s = h->stream;
    const size_t iq_bytes = (size_t)iq_samples * sample_bytes(h->fmt);
    if (iq_bytes > h->hs[0].d_iq_cap) {
        if (h->hs[0].d_iq) { CU(h, cudaStreamSynchronize(s)); cudaFree(h->hs[0].d_iq); h->hs[0].d_iq = nullptr; h->hs[0].d_iq_cap = 0; }
        cudaError_t e = cudaMalloc((void **)&h->hs[0].d_iq, iq_bytes);
        if (e != cudaSuccess) return fail(h, B200RX_E_NOMEM, "b200rx_decode_headers: sample staging", e);
        h->hs[0].d_iq_cap = iq_bytes;
    }
    CU(h, cudaMemcpyAsync(h->hs[0].d_iq, iq, iq_bytes, cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->hs[0].d_lts1, lts1_index, n_frames * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->hs[0].d_avail, avail, n_frames * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    FrontendArgs fa{};
    fa.iq = h->hs[0].d_iq;
    fa.fmt = h->fmt;
    fa.scale = h->scale;
    fa.iq_samples = iq_samples;
    fa.lts1 = h->hs[0].d_lts1;
    fa.avail = h->hs[0].d_avail;
    fa.n_frames = n_frames;
    fa.desc = h->desc;
    fa.bm = h->bm;
    fa.bm_stride = h->max_steps;
    fa.max_steps = h->max_steps;
    fa.max_len = h->limits.max_payload_bytes;
    fa.header_only = 1;
    CU(h, launch_frontend(fa, s));
    CU(h, launch_export_headers(h->desc, n_frames, h->hs[0].d_len, h->hs[0].d_rate, h->hs[0].d_status, s));
    h->launches += 2;
    if (payload_len) CU(h, cudaMemcpyAsync(payload_len, h->hs[0].d_len, n_frames * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
    if (rate_out) CU(h, cudaMemcpyAsync(rate_out, h->hs[0].d_rate, n_frames, cudaMemcpyDeviceToHost, s));
    CU(h, cudaMemcpyAsync(status, h->hs[0].d_status, n_frames, cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    return B200RX_OK;
}

int b200rx_viterbi_batch_dev(b200rx_handle *h, const uint8_t *symbols_dev, uint64_t symbols_stride,
                             const uint32_t *data_bits_dev, uint32_t max_data_bits, uint32_t n_frames,
                             uint8_t *out_dev, uint32_t out_stride)
{
    if (!h) return B200RX_E_ARG;
    if (!symbols_dev || !data_bits_dev || !out_dev) return fail(h, B200RX_E_ARG, "b200rx_viterbi_batch_dev: null argument");
    if (n_frames > h->limits.max_frames) return fail(h, B200RX_E_ARG, "b200rx_viterbi_batch_dev: n_frames exceeds max_frames");
    if (max_data_bits + 6 > h->max_steps) return fail(h, B200RX_E_ARG, "b200rx_viterbi_batch_dev: trellis longer than the handle's capacity");
    if ((symbols_stride & 1) || ((uintptr_t)symbols_dev & 1)) return fail(h, B200RX_E_ARG, "b200rx_viterbi_batch_dev: symbols must be 2-byte aligned");
    if (n_frames == 0) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    { int rcq = drain_lanes(h); if (rcq != B200RX_OK) return rcq; } // not pipelined: work on scratch set 0
    cudaStream_t s = h->stream;
    CU(h, cudaMemsetAsync(h->counters, 0, 8 * sizeof(unsigned long long), s));
    CU(h, cudaEventRecord(h->ev[0], s));
    if (h->tn.acs_gen == 3) { // the ACS kernel reads the caller's symbols in place
        CU(h, launch_desc_from_bits(data_bits_dev, n_frames, h->desc, h->max_steps, s));
        CU(h, cudaEventRecord(h->ev[1], s));
        CU(h, launch_viterbi_acs3(h->desc, symbols_dev, symbols_stride, h->dec, 2 * h->max_steps, n_frames, true, h->tn, s));
    } else {
        CU(h, launch_bm_from_symbols(symbols_dev, symbols_stride, data_bits_dev, max_data_bits, n_frames, h->desc, h->bm,
                                     h->max_steps, h->max_steps, s));
        CU(h, cudaEventRecord(h->ev[1], s));
        CU(h, launch_viterbi_acs(h->desc, h->bm, h->max_steps, h->dec, 2 * h->max_steps, n_frames, h->tn, s));
    }
    CU(h, cudaEventRecord(h->ev[2], s));
    TracebackArgs ta{};
    ta.desc = h->desc;
    ta.dec = h->dec;
    ta.dec_stride = 2 * h->max_steps;
    ta.n_frames = n_frames;
    ta.raw_mode = 1;
    ta.payload = out_dev;
    ta.payload_stride = out_stride;
    ta.counters = h->counters;
    CU(h, launch_traceback(ta, h->tn, s));
    CU(h, cudaEventRecord(h->ev[3], s));
    h->ev_valid = true;
    h->launches += 3;
    return B200RX_OK;
}

int b200rx_profile_begin(b200rx_handle *h, uint32_t slots)
try {
    if (!h) return B200RX_E_ARG;
    if (slots > 65536) return fail(h, B200RX_E_ARG, "b200rx_profile_begin: at most 65536 slots");
    CU(h, cudaSetDevice(h->device));
    for (cudaEvent_t e : h->ring) if (e) cudaEventDestroy(e);
    h->ring.assign(4 * (size_t)slots, nullptr);
    for (size_t i = 0; i < h->ring.size(); i++) CU(h, cudaEventCreate(&h->ring[i]));
    h->ring_slots = slots;
    h->ring_used = 0;
    return B200RX_OK;
}
catch (...) { return B200RX_E_NOMEM; } // std::bad_alloc must not cross the C boundary

int b200rx_profile_read(b200rx_handle *h, uint32_t *calls, float *frontend_ms, float *viterbi_ms, float *traceback_ms)
{
    if (!h || !calls || !frontend_ms || !viterbi_ms || !traceback_ms) return B200RX_E_ARG;
    CU(h, cudaSetDevice(h->device));
    { int rcq = b200rx_synchronize(h); if (rcq != B200RX_OK) return rcq; }
    *calls = h->ring_used;
    *frontend_ms = *viterbi_ms = *traceback_ms = 0.f;
    for (uint32_t i = 0; i < h->ring_used; i++) {
        cudaEvent_t *e = &h->ring[4 * (size_t)i];
        float a, b, c;
        CU(h, cudaEventElapsedTime(&a, e[0], e[1]));
        CU(h, cudaEventElapsedTime(&b, e[1], e[2]));
        CU(h, cudaEventElapsedTime(&c, e[2], e[3]));
        *frontend_ms += a; *viterbi_ms += b; *traceback_ms += c;
    }
    h->ring_slots = 0; // profiling ends with the read
    return B200RX_OK;
}

int b200rx_get_stats(b200rx_handle *h, b200rx_stats *out)
{
    if (!h || !out) return B200RX_E_ARG;
    memset(out, 0, sizeof(*out));
    if (!h->ev_valid) return B200RX_OK;
    CU(h, cudaSetDevice(h->device));
    { int rcq = b200rx_synchronize(h); if (rcq != B200RX_OK) return rcq; }
    CU(h, cudaEventElapsedTime(&out->frontend_ms, h->ev[0], h->ev[1]));
    CU(h, cudaEventElapsedTime(&out->viterbi_ms, h->ev[1], h->ev[2]));
    CU(h, cudaEventElapsedTime(&out->traceback_ms, h->ev[2], h->ev[3]));
    CU(h, cudaEventElapsedTime(&out->total_ms, h->ev[0], h->ev[3]));
    unsigned long long c[8];
    CU(h, cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
    out->traceback_rewalks = c[4];
    out->frames_ok = (uint32_t)c[0];
    out->frames_failed = (uint32_t)c[1];
    out->payload_bytes = c[2];
    out->trellis_steps = c[3];
    return B200RX_OK;
}

} // extern "C"
