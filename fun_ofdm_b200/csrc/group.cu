// Several GPUs of one box from one process (include/b200rx.h: b200rx_group_*, b200rx_gather_status).
// Written against the public C ABI only: a group is n handles + n host threads; the data path has no collective.
// NCCL (status gather + counter sum) is loaded with dlopen on first use, so the library itself does not depend on it.
#include "../../include/b200rx.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace {

thread_local char g_group_error[512] = {0};

// ---- the few NCCL entry points used, resolved at run time (prototypes as in nccl.h 2.x) ----
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t; // ncclSuccess == 0
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_SUM = 0 }; // ncclDataType_t / ncclRedOp_t values (stable since NCCL 2.0)
struct NcclApi {
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!so) return;
        api.CommInitAll = (decltype(api.CommInitAll))dlsym(so, "ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(so, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(so, "ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))dlsym(so, "ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))dlsym(so, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(so, "ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(so, "ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.AllGather && api.AllReduce && api.GroupStart && api.GroupEnd;
    });
    return api;
}

// One worker per device: runs the jobs handed to it in order, on a thread that has made the device current once.
class Worker {
public:
    explicit Worker(int device) : m_device(device), m_thread([this] { loop(); }) {}
    ~Worker()
    {
        {
            std::lock_guard<std::mutex> l(m_mu);
            m_stop = true;
        }
        m_cv.notify_all();
        m_thread.join();
    }
    void post(std::function<int()> job)
    {
        {
            std::lock_guard<std::mutex> l(m_mu);
            m_job = std::move(job);
            m_has_job = true;
            m_done = false;
        }
        m_cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> l(m_mu);
        m_cv.wait(l, [this] { return m_done; });
        return m_rc;
    }

private:
    void loop()
    {
        cudaSetDevice(m_device);
        for (;;) {
            std::function<int()> job;
            {
                std::unique_lock<std::mutex> l(m_mu);
                m_cv.wait(l, [this] { return m_has_job || m_stop; });
                if (m_stop) return;
                job.swap(m_job);
                m_has_job = false;
            }
            int rc = B200RX_E_NOMEM;
            try { rc = job(); } catch (...) { rc = B200RX_E_NOMEM; }
            {
                std::lock_guard<std::mutex> l(m_mu);
                m_rc = rc;
                m_done = true;
            }
            m_cv.notify_all();
        }
    }
    int m_device;
    std::mutex m_mu;
    std::condition_variable m_cv;
    std::function<int()> m_job;
    bool m_has_job = false, m_done = true, m_stop = false;
    int m_rc = 0;
    std::thread m_thread; // last: starts when everything above exists
};

} // namespace

struct b200rx_group {
    std::vector<int> devices;
    std::vector<b200rx_handle *> handles;
    std::vector<Worker *> workers;
    std::vector<cudaStream_t> comm_stream;
    std::vector<unsigned long long *> sum_dev; // [4] per device: all-reduced counters
    std::vector<ncclComm_t> comms;
    unsigned long long *sum_host = nullptr;    // pinned [4]
    b200rx_limits limits{};
    char error_buf[512] = {0};
};

namespace {

int gfail(b200rx_group *g, int code, const char *what, const char *detail = nullptr)
{
    char *buf = g ? g->error_buf : g_group_error;
    if (detail) snprintf(buf, 512, "%s: %s", what, detail);
    else snprintf(buf, 512, "%s", what);
    return code;
}

// run job(i) on every device's thread; first failure wins
int fan_out(b200rx_group *g, const std::function<int(uint32_t)> &job, const char *what)
{
    const uint32_t n = (uint32_t)g->handles.size();
    for (uint32_t i = 0; i < n; i++) g->workers[i]->post([&job, i] { return job(i); });
    int rc = B200RX_OK;
    for (uint32_t i = 0; i < n; i++) {
        const int r = g->workers[i]->wait();
        if (r != B200RX_OK && rc == B200RX_OK) {
            rc = r;
            char d[400];
            snprintf(d, sizeof(d), "device %d: %s", g->devices[i], b200rx_last_error(g->handles[i]));
            gfail(g, r, what, d);
        }
    }
    return rc;
}

int ensure_comms(b200rx_group *g)
{
    if (!g->comms.empty()) return B200RX_OK;
    NcclApi &api = nccl();
    if (!api.ok) return gfail(g, B200RX_E_DEVICE, "b200rx_gather_status: libnccl.so.2 not found");
    const int n = (int)g->devices.size();
    g->comms.assign(n, nullptr);
    ncclResult_t r = api.CommInitAll(g->comms.data(), n, g->devices.data());
    if (r != 0) {
        g->comms.clear();
        return gfail(g, B200RX_E_CUDA, "ncclCommInitAll", api.GetErrorString ? api.GetErrorString(r) : "failed");
    }
    return B200RX_OK;
}

} // namespace

extern "C" {

const char *b200rx_group_last_error(const b200rx_group *g) { return g ? g->error_buf : g_group_error; }

uint32_t b200rx_group_size(const b200rx_group *g) { return g ? (uint32_t)g->handles.size() : 0; }

b200rx_handle *b200rx_group_handle(b200rx_group *g, uint32_t i) { return (g && i < g->handles.size()) ? g->handles[i] : nullptr; }

int b200rx_group_create(const int *devices, uint32_t n_devices, const b200rx_limits *limits, b200rx_group **out)
{
    if (!out || !devices || !limits || n_devices == 0 || n_devices > 64)
        return gfail(nullptr, B200RX_E_ARG, "b200rx_group_create: bad argument");
    *out = nullptr;
    for (uint32_t i = 0; i < n_devices; i++)
        for (uint32_t k = 0; k < i; k++)
            if (devices[i] == devices[k]) return gfail(nullptr, B200RX_E_ARG, "b200rx_group_create: a device is listed twice");
    b200rx_group *g = nullptr;
    try {
        g = new b200rx_group();
        g->devices.assign(devices, devices + n_devices);
        g->limits = *limits;
        g->handles.assign(n_devices, nullptr);
        g->comm_stream.assign(n_devices, nullptr);
        g->sum_dev.assign(n_devices, nullptr);
        for (uint32_t i = 0; i < n_devices; i++) g->workers.push_back(new Worker(devices[i]));
    } catch (...) {
        if (g) b200rx_group_destroy(g);
        return gfail(nullptr, B200RX_E_NOMEM, "b200rx_group_create: out of host memory");
    }
    // every handle is created on its own thread (scratch allocation of the devices proceeds in parallel)
    std::vector<std::vector<char>> err(n_devices, std::vector<char>(512, 0));
    int rc = B200RX_OK;
    for (uint32_t i = 0; i < n_devices; i++)
        g->workers[i]->post([g, i, &err] {
            int r = b200rx_create(g->devices[i], &g->limits, &g->handles[i]);
            if (r != B200RX_OK) { snprintf(err[i].data(), 512, "%s", b200rx_last_error(nullptr)); return r; }
            if (cudaStreamCreateWithFlags(&g->comm_stream[i], cudaStreamNonBlocking) != cudaSuccess) return B200RX_E_CUDA;
            if (cudaMalloc((void **)&g->sum_dev[i], 4 * sizeof(unsigned long long)) != cudaSuccess) return B200RX_E_NOMEM;
            return B200RX_OK;
        });
    for (uint32_t i = 0; i < n_devices; i++) {
        const int r = g->workers[i]->wait();
        if (r != B200RX_OK && rc == B200RX_OK) {
            rc = r;
            char d[400];
            snprintf(d, sizeof(d), "device %d: %s", devices[i], err[i].data());
            gfail(nullptr, r, "b200rx_group_create", d);
        }
    }
    if (rc == B200RX_OK && cudaHostAlloc((void **)&g->sum_host, 4 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess)
        rc = gfail(nullptr, B200RX_E_NOMEM, "b200rx_group_create: pinned counters");
    if (rc != B200RX_OK) {
        b200rx_group_destroy(g);
        return rc;
    }
    *out = g;
    return B200RX_OK;
}

int b200rx_group_destroy(b200rx_group *g)
{
    if (!g) return B200RX_OK;
    for (size_t i = 0; i < g->workers.size(); i++) {
        Worker *w = g->workers[i];
        if (!w) continue;
        w->post([g, i] {
            if (i < g->comms.size() && g->comms[i] && nccl().ok) nccl().CommDestroy(g->comms[i]);
            if (g->handles[i]) b200rx_destroy(g->handles[i]);
            if (g->comm_stream[i]) cudaStreamDestroy(g->comm_stream[i]);
            cudaFree(g->sum_dev[i]);
            return B200RX_OK;
        });
        w->wait();
        delete w;
    }
    if (g->sum_host) cudaFreeHost(g->sum_host);
    delete g;
    return B200RX_OK;
}

int b200rx_group_synchronize(b200rx_group *g)
{
    if (!g) return B200RX_E_ARG;
    return fan_out(g, [g](uint32_t i) { return b200rx_synchronize(g->handles[i]); }, "b200rx_group_synchronize");
}

int b200rx_group_plan(const b200rx_group *g, const uint32_t *weight, uint32_t n_frames, uint32_t *first)
{
    if (!g || !first) return B200RX_E_ARG;
    const uint32_t n = (uint32_t)g->handles.size();
    unsigned long long total = 0;
    for (uint32_t f = 0; f < n_frames; f++) total += weight ? (unsigned long long)weight[f] + 1ull : 1ull;
    // device i ends where the running weight first reaches (i + 1) / n of the total
    unsigned long long run = 0;
    uint32_t f = 0;
    first[0] = 0;
    for (uint32_t i = 0; i < n; i++) {
        const unsigned long long target = total * (i + 1) / n;
        while (f < n_frames && run < target) {
            run += weight ? (unsigned long long)weight[f] + 1ull : 1ull;
            f++;
        }
        first[i + 1] = (i + 1 == n) ? n_frames : f;
    }
    return B200RX_OK;
}

int b200rx_group_decode_batch(b200rx_group *g, const void *iq, uint64_t iq_samples, const uint64_t *lts1_index,
                              const uint32_t *avail, uint32_t n_frames, uint8_t *payload_out, uint32_t payload_stride,
                              uint16_t *payload_len, uint8_t *rate_out, uint8_t *status)
{
    if (!g) return B200RX_E_ARG;
    if (!iq || !lts1_index || !avail || !status) return gfail(g, B200RX_E_ARG, "b200rx_group_decode_batch: null argument");
    if (n_frames == 0) return B200RX_OK;
    const uint32_t n = (uint32_t)g->handles.size();
    std::vector<uint32_t> first;
    std::vector<std::vector<uint64_t>> rebased;
    try {
        first.resize(n + 1);
        rebased.resize(n);
    } catch (...) { return gfail(g, B200RX_E_NOMEM, "b200rx_group_decode_batch: out of host memory"); }
    b200rx_group_plan(g, avail, n_frames, first.data());
    for (uint32_t i = 0; i < n; i++)
        if (first[i + 1] - first[i] > g->limits.max_frames)
            return gfail(g, B200RX_E_ARG, "b200rx_group_decode_batch: a shard exceeds max_frames");
    return fan_out(g, [&](uint32_t i) -> int {
        const uint32_t f0 = first[i], nf = first[i + 1] - f0;
        if (nf == 0) return B200RX_OK;
        // the shard's own window of the sample buffer: a device stages only what its frames touch
        uint64_t lo = UINT64_MAX, hi = 0;
        for (uint32_t f = f0; f < f0 + nf; f++) {
            const uint64_t a = lts1_index[f] < iq_samples ? lts1_index[f] : iq_samples;
            uint64_t e = a + avail[f];
            if (e > iq_samples) e = iq_samples;
            if (a < lo) lo = a;
            if (e > hi) hi = e;
        }
        if (hi < lo) hi = lo;
        rebased[i].resize(nf);
        for (uint32_t f = 0; f < nf; f++) {
            const uint64_t a = lts1_index[f0 + f];
            rebased[i][f] = a >= lo ? a - lo : (uint64_t)-1; // beyond the buffer: stays out of range -> TRUNCATED
            if (a >= iq_samples) rebased[i][f] = hi - lo;
        }
        const size_t bps = b200rx_sample_bytes(g->handles[i]);
        return b200rx_decode_batch(g->handles[i], (const uint8_t *)iq + lo * bps, hi - lo, rebased[i].data(), avail + f0, nf,
                                   payload_out ? payload_out + (size_t)f0 * payload_stride : nullptr, payload_stride,
                                   payload_len ? payload_len + f0 : nullptr, rate_out ? rate_out + f0 : nullptr, status + f0);
    }, "b200rx_group_decode_batch");
}

int b200rx_group_decode_batch_dev(b200rx_group *g, const void *const *iq_dev, const uint64_t *iq_samples,
                                  const uint64_t *const *lts1_index_dev, const uint32_t *const *avail_dev,
                                  const uint32_t *n_frames, uint8_t *const *payload_out_dev, uint32_t payload_stride,
                                  uint16_t *const *payload_len_dev, uint8_t *const *rate_out_dev, uint8_t *const *status_dev)
{
    if (!g) return B200RX_E_ARG;
    if (!iq_dev || !iq_samples || !lts1_index_dev || !avail_dev || !n_frames || !status_dev)
        return gfail(g, B200RX_E_ARG, "b200rx_group_decode_batch_dev: null argument");
    return fan_out(g, [&](uint32_t i) -> int {
        if (n_frames[i] == 0) return B200RX_OK;
        return b200rx_decode_batch_dev(g->handles[i], iq_dev[i], iq_samples[i], lts1_index_dev[i], avail_dev[i], n_frames[i],
                                       payload_out_dev ? payload_out_dev[i] : nullptr, payload_stride,
                                       payload_len_dev ? payload_len_dev[i] : nullptr, rate_out_dev ? rate_out_dev[i] : nullptr,
                                       status_dev[i], nullptr);
    }, "b200rx_group_decode_batch_dev");
}

int b200rx_gather_status(b200rx_group *g, const uint8_t *const *status_dev, uint32_t frames_per_device,
                         uint8_t *const *gathered_dev, uint64_t *counters_sum)
try {
    if (!g) return B200RX_E_ARG;
    int rc = ensure_comms(g);
    if (rc != B200RX_OK) return rc;
    NcclApi &api = nccl();
    const uint32_t n = (uint32_t)g->handles.size();
    const bool gather = status_dev && gathered_dev && frames_per_device;
    // each device's communication stream waits for everything its handle has been given so far
    std::vector<void *> ctr(n, nullptr);
    for (uint32_t i = 0; i < n; i++) {
        if (cudaSetDevice(g->devices[i]) != cudaSuccess) return gfail(g, B200RX_E_CUDA, "b200rx_gather_status: cudaSetDevice");
        rc = b200rx_join_on(g->handles[i], 0, g->comm_stream[i]);
        if (rc == B200RX_OK) rc = b200rx_device_counters(g->handles[i], &ctr[i]);
        if (rc != B200RX_OK) return gfail(g, rc, "b200rx_gather_status", b200rx_last_error(g->handles[i]));
    }
    ncclResult_t r = api.GroupStart();
    for (uint32_t i = 0; i < n && r == 0; i++) {
        if (gather) r = api.AllGather(status_dev[i], gathered_dev[i], frames_per_device, NCCL_UINT8, g->comms[i], g->comm_stream[i]);
        if (r == 0) r = api.AllReduce(ctr[i], g->sum_dev[i], 4, NCCL_UINT64, NCCL_SUM, g->comms[i], g->comm_stream[i]);
    }
    const ncclResult_t r2 = api.GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) return gfail(g, B200RX_E_CUDA, "b200rx_gather_status: NCCL", api.GetErrorString ? api.GetErrorString(r) : "failed");
    for (uint32_t i = 0; i < n; i++) {
        if (cudaSetDevice(g->devices[i]) != cudaSuccess) return gfail(g, B200RX_E_CUDA, "b200rx_gather_status: cudaSetDevice");
        if (i == 0 && counters_sum)
            if (cudaMemcpyAsync(g->sum_host, g->sum_dev[0], 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                g->comm_stream[0]) != cudaSuccess)
                return gfail(g, B200RX_E_CUDA, "b200rx_gather_status: counters to host");
        if (cudaStreamSynchronize(g->comm_stream[i]) != cudaSuccess)
            return gfail(g, B200RX_E_CUDA, "b200rx_gather_status: synchronising a communication stream");
    }
    if (counters_sum)
        for (int k = 0; k < 4; k++) counters_sum[k] = g->sum_host[k];
    return B200RX_OK;
}
catch (...) { return B200RX_E_NOMEM; }

} // extern "C"
