// TEST DOUBLE — not product code.  Implements the part of include/b200rx.h that the host adapters
// (fun::b200_rx, fun::b200_receiver_chain) call, on the CPU, by running the UNMODIFIED reference blocks of
// oracle/_ref/libfunref.so (dlopen).  It exists so that the adapters' bookkeeping - frame cutting, streaming state,
// chunk boundaries, deduplication - is exercised by the CPU test stage; the GPU tests exercise the same adapters
// against the real library.  Never linked into anything under fun_ofdm_b200/.
#include "../../include/b200rx.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

namespace {

struct ref_frame_info { int32_t hdr_ok, hdr_field, hdr_parity, rate_valid, rate, length, nsym, crc_ok, n_vectors, payload_from_blocks; };
typedef int (*decode_frame_fn)(const double *, int, ref_frame_info *, double *, int, uint8_t *, uint8_t *, uint8_t *, uint8_t *,
                               uint8_t *, uint8_t *);
typedef long (*sync_bounds_fn)(const double *, long, const long *, int, int, double *, uint8_t *);

decode_frame_fn g_decode = nullptr;
sync_bounds_fn g_sync = nullptr;
std::string g_err;

bool load_ref()
{
    if (g_decode && g_sync) return true;
    const char *path = getenv("B200RX_FAKE_REF");
    void *so = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
    if (!so) { g_err = "fake b200rx: cannot load B200RX_FAKE_REF"; return false; }
    g_decode = (decode_frame_fn)dlsym(so, "ref_decode_frame");
    g_sync = (sync_bounds_fn)dlsym(so, "ref_sync_bounds");
    if (!g_decode || !g_sync) { g_err = "fake b200rx: reference harness symbols missing"; return false; }
    return true;
}

// one window (samples from the LTS1 tag on) through the reference's four blocks -> what the C ABI reports
void decode_window(const double *iq, uint32_t n, uint32_t max_len, uint8_t *payload, uint16_t *len, uint8_t *rate, uint8_t *status,
                   bool header_only)
{
    ref_frame_info info;
    std::vector<uint8_t> pl(4096 + 16);
    g_decode(iq, (int)n, &info, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, pl.data());
    *len = 0; *rate = B200RX_RATE_INVALID;
    if (n < 208 || info.n_vectors < 1) { *status = B200RX_ST_TRUNCATED; return; }
    if (!info.hdr_ok) { *status = info.hdr_parity ? B200RX_ST_HDR_PARITY : B200RX_ST_HDR_RATE; return; }
    *len = (uint16_t)info.length; *rate = (uint8_t)info.rate;
    if ((uint32_t)info.length > max_len) { *status = B200RX_ST_TOO_LONG; return; }
    if (header_only) { *status = B200RX_ST_OK; return; }
    if (info.n_vectors < 1 + info.nsym) { *status = B200RX_ST_TRUNCATED; return; }
    *status = info.crc_ok ? B200RX_ST_OK : B200RX_ST_CRC_FAIL;
    if (info.crc_ok && payload && info.length) memcpy(payload, pl.data(), (size_t)info.length);
}

} // namespace

struct b200rx_handle {
    b200rx_limits lim;
    std::vector<int64_t> origins;
    std::string error;
    // two-phase passes: the staged capture and what the reference made of it
    std::vector<double> pass_iq;
    std::vector<uint8_t> pass_raw; // tagged passes: the 24-byte structs as put
    int fmt = B200RX_FMT_FC64;
    std::vector<uint8_t> pass_payload, pass_status;
    std::vector<uint16_t> pass_len;
    uint32_t pass_frames = 0;
};

extern "C" {

int b200rx_create(int, const b200rx_limits *limits, b200rx_handle **out)
{
    if (!load_ref()) return B200RX_E_DEVICE;
    *out = new b200rx_handle();
    (*out)->lim = *limits;
    return B200RX_OK;
}
int b200rx_destroy(b200rx_handle *h) { delete h; return B200RX_OK; }
const char *b200rx_last_error(const b200rx_handle *h) { return h ? h->error.c_str() : g_err.c_str(); }
const char *b200rx_version(void) { return "fake b200rx (CPU test double over oracle/_ref)"; }
// "pinned" memory of the double: plain malloc, remembered so that b200rx_host_is_pinned can answer like the real library
// (the chain copies long calls to the GPU straight from pinned caller buffers: that path is exercised on the CPU too)
static std::vector<std::pair<char *, size_t> > g_pinned;
int b200rx_host_alloc(void **p, size_t bytes)
{
    *p = malloc(bytes ? bytes : 1);
    if (*p) g_pinned.push_back(std::make_pair((char *)*p, bytes ? bytes : 1));
    return *p ? B200RX_OK : B200RX_E_NOMEM;
}
int b200rx_host_free(void *p)
{
    for (size_t i = 0; i < g_pinned.size(); i++)
        if (g_pinned[i].first == (char *)p) { g_pinned.erase(g_pinned.begin() + i); break; }
    free(p);
    return B200RX_OK;
}

int b200rx_set_receive_origins(b200rx_handle *h, const int64_t *origins, uint32_t n)
{
    h->origins.assign(origins, origins + n);
    return B200RX_OK;
}

int b200rx_decode_headers(b200rx_handle *h, const void *iq, uint64_t, const uint64_t *lts1, const uint32_t *avail, uint32_t n,
                          uint16_t *len, uint8_t *rate, uint8_t *status)
{
    for (uint32_t f = 0; f < n; f++)
        decode_window((const double *)iq + 2 * lts1[f], avail[f], h->lim.max_payload_bytes, nullptr, len + f, rate + f, status + f, true);
    return B200RX_OK;
}

int b200rx_decode_batch(b200rx_handle *h, const void *iq, uint64_t, const uint64_t *lts1, const uint32_t *avail, uint32_t n,
                        uint8_t *payload, uint32_t stride, uint16_t *len, uint8_t *rate, uint8_t *status)
{
    for (uint32_t f = 0; f < n; f++)
        decode_window((const double *)iq + 2 * lts1[f], avail[f], h->lim.max_payload_bytes, payload + (size_t)f * stride, len + f,
                      rate + f, status + f, false);
    return B200RX_OK;
}

// One capture through the reference's frame_detector + timing_sync (work() calls cut where the origins say the
// reference's own calls were cut), then every LTS1-tagged window through its four hot-path blocks.
int b200rx_receive(b200rx_handle *h, const void *iq, uint64_t n, double, uint8_t *payload, uint32_t stride, uint16_t *len,
                   uint8_t *rate, uint8_t *status, uint64_t *lts1_out, b200rx_sync_result *res)
{
    memset(res, 0, sizeof(*res));
    std::vector<long> bounds(1, 0L);
    for (int64_t o : h->origins) {
        const int64_t c = o + 160; // chunk start
        if (c > bounds.back() && c < (int64_t)n) bounds.push_back((long)c);
    }
    bounds.push_back((long)n);
    h->origins.clear();
    if (n == 0) return B200RX_OK;
    std::vector<double> out(2 * (n + 160));
    std::vector<uint8_t> tags(n + 160);
    const long got = g_sync((const double *)iq, (long)n, bounds.data(), (int)bounds.size(), 160, out.data(), tags.data());
    // output index j carries stream sample j - 160; STS_END tags in the last 160 samples wait for the next capture
    // (timing_sync.cpp:68), which here means: LTS1 tags set by events at x >= n - 160 are ignored
    std::vector<uint64_t> starts;
    int64_t last_event = -1;
    for (long j = 160; j < got; j++) {
        const uint64_t p = (uint64_t)(j - 160);
        if (tags[j] == 2) last_event = (int64_t)p;        // STS_END
        if (tags[j] == 4 && p < n) {                      // LTS1
            if (last_event >= (int64_t)n - 160) continue;  // found by an event the GPU path defers
            starts.push_back(p);
        }
    }
    uint32_t nf = 0;
    for (size_t k = 0; k < starts.size() && nf < h->lim.max_frames; k++, nf++) {
        const uint64_t a = starts[k], e = k + 1 < starts.size() ? starts[k + 1] : n;
        decode_window(out.data() + 2 * (a + 160), (uint32_t)(e - a), h->lim.max_payload_bytes, payload + (size_t)nf * stride,
                      len + nf, rate + nf, status + nf, false);
        if (lts1_out) lts1_out[nf] = a;
    }
    res->n_frames = nf;
    res->overflow = (uint32_t)(starts.size() - nf);
    return B200RX_OK;
}


// ---- two-phase passes: scan = the same capture logic, decode = hand out what the scan already computed ----
int b200rx_set_pipeline_depth(b200rx_handle *, uint32_t) { return B200RX_OK; }
int b200rx_set_tuning(b200rx_handle *, const char *, int64_t) { return B200RX_OK; }
int b200rx_host_is_pinned(const void *p)
{
    for (size_t i = 0; i < g_pinned.size(); i++)
        if ((const char *)p >= g_pinned[i].first && (const char *)p < g_pinned[i].first + g_pinned[i].second) return 1;
    return 0;
}
int b200rx_set_sample_format(b200rx_handle *h, int fmt, double) { h->fmt = fmt; return B200RX_OK; }
int b200rx_pass_open(b200rx_handle *h) { h->pass_iq.clear(); h->pass_raw.clear(); h->pass_frames = 0; return B200RX_OK; }
int b200rx_pass_put(b200rx_handle *h, const void *iq, uint64_t n)
{
    if (h->fmt == B200RX_FMT_TAGGED_FC64) {
        const uint8_t *p = (const uint8_t *)iq;
        h->pass_raw.insert(h->pass_raw.end(), p, p + 24 * n);
        return B200RX_OK;
    }
    const double *p = (const double *)iq;
    h->pass_iq.insert(h->pass_iq.end(), p, p + 2 * n);
    return B200RX_OK;
}
// a stream timing_sync has tagged: frames start at the LTS1 tags, each window through the reference's four blocks
int b200rx_pass_scan_tagged(b200rx_handle *h, b200rx_pass_frame *frames, uint32_t cap, b200rx_sync_result *res)
{
    memset(res, 0, sizeof(*res));
    const uint32_t mf = h->lim.max_frames, stride = h->lim.max_payload_bytes ? h->lim.max_payload_bytes : 1;
    h->pass_payload.assign((size_t)mf * stride, 0);
    h->pass_status.assign(mf, B200RX_ST_NO_FRAME);
    h->pass_len.assign(mf, 0);
    const uint64_t n = h->pass_raw.size() / 24;
    std::vector<double> iq(2 * n);
    std::vector<uint64_t> starts;
    for (uint64_t i = 0; i < n; i++) {
        memcpy(&iq[2 * i], &h->pass_raw[24 * i], 16);
        int32_t tag;
        memcpy(&tag, &h->pass_raw[24 * i + 16], 4);
        if (tag == 4) starts.push_back(i);
    }
    uint32_t nf = 0;
    for (size_t k = 0; k < starts.size() && nf < mf; k++, nf++) {
        const uint64_t a = starts[k], e = k + 1 < starts.size() ? starts[k + 1] : n;
        uint8_t rate = 0;
        decode_window(iq.data() + 2 * a, (uint32_t)(e - a), h->lim.max_payload_bytes, h->pass_payload.data() + (size_t)nf * stride,
                      &h->pass_len[nf], &rate, &h->pass_status[nf], false);
        if (nf < cap) {
            frames[nf].lts1 = a;
            frames[nf].avail = (uint32_t)(e - a);
            frames[nf].length = h->pass_len[nf];
            frames[nf].rate = rate;
            frames[nf].status = h->pass_status[nf] == B200RX_ST_CRC_FAIL ? (uint8_t)B200RX_ST_OK : h->pass_status[nf];
            frames[nf].sts_end = 0;
            frames[nf].phase = 0.0;
        }
    }
    h->pass_frames = nf;
    res->n_events = (uint32_t)starts.size();
    res->n_frames = nf;
    res->overflow = (uint32_t)(starts.size() - nf);
    return B200RX_OK;
}
int b200rx_pass_scan(b200rx_handle *h, double phase_in, b200rx_pass_frame *frames, uint32_t cap, b200rx_sync_result *res)
{
    const uint32_t mf = h->lim.max_frames, stride = h->lim.max_payload_bytes ? h->lim.max_payload_bytes : 1;
    h->pass_payload.assign((size_t)mf * stride, 0);
    h->pass_status.assign(mf, B200RX_ST_NO_FRAME);
    h->pass_len.assign(mf, 0);
    std::vector<uint8_t> rate(mf);
    std::vector<uint64_t> lts1(mf);
    const uint64_t n = h->pass_iq.size() / 2;
    int rc = b200rx_receive(h, h->pass_iq.data(), n, phase_in, h->pass_payload.data(), stride, h->pass_len.data(), rate.data(),
                            h->pass_status.data(), lts1.data(), res);
    if (rc != B200RX_OK) return rc;
    h->pass_frames = res->n_frames;
    for (uint32_t f = 0; f < res->n_frames && f < cap; f++) {
        frames[f].lts1 = lts1[f];
        frames[f].avail = (uint32_t)((f + 1 < res->n_frames ? lts1[f + 1] : n) - lts1[f]);
        frames[f].length = h->pass_len[f];
        frames[f].rate = rate[f];
        const uint8_t st = h->pass_status[f];
        frames[f].status = (st == B200RX_ST_CRC_FAIL) ? (uint8_t)B200RX_ST_OK : st; // the header verdict only
        frames[f].sts_end = 0;  // the double carries no phase state: its reference blocks synchronise every capture afresh
        frames[f].phase = 0.0;
    }
    return B200RX_OK;
}
int b200rx_pass_decode(b200rx_handle *h, const uint8_t *select, uint8_t *payload, uint32_t stride, uint8_t *status, uint64_t *ticket)
{
    const uint32_t own = h->lim.max_payload_bytes ? h->lim.max_payload_bytes : 1;
    static uint64_t next_ticket = 1;
    *ticket = 0;
    for (uint32_t f = 0; f < h->pass_frames; f++) {
        if (!select[f]) continue;
        if (*ticket == 0) *ticket = next_ticket++;
        status[f] = h->pass_status[f];
        if (payload && status[f] == B200RX_ST_OK)
            memcpy(payload + (size_t)f * stride, h->pass_payload.data() + (size_t)f * own, h->pass_len[f] < stride ? h->pass_len[f] : stride);
    }
    return B200RX_OK;
}
// B200RX_FAKE_SLOW_POLLS=k: a pass reports "still running" to its first k polls (a wait always completes it) - lets the
// CPU tests see the adapters' delivery-lag rules, which a double that finishes instantly would never exercise
static int g_polls_left = -1;
int b200rx_pass_poll(b200rx_handle *, uint64_t ticket)
{
    if (ticket == 0) return 1;
    static uint64_t cur = 0;
    static int left = 0;
    if (g_polls_left < 0) { const char *e = getenv("B200RX_FAKE_SLOW_POLLS"); g_polls_left = e ? atoi(e) : 0; }
    if (ticket != cur) { cur = ticket; left = g_polls_left; }
    if (left > 0) { left--; return 0; }
    return 1;
}
int b200rx_pass_wait(b200rx_handle *, uint64_t) { return B200RX_OK; }

} // extern "C"
