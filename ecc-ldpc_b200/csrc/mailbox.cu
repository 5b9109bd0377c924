// mailbox.cu -- the reference tester's calling pattern made fast: ONE frame per `decode` call, from several host threads at
// once, each with its own decoder handle (src/ECC/Code/LDPC/Utils.hs:53,62-69: `decoder's !! (tid mod maxThreadCount)`).
//
// Through the batched path such a call costs an H2D copy, a conversion pass, the decode kernel, an unpack pass, a D2H copy
// and a stream synchronisation: ~80 us, of which the decode itself is less than half, and every thread pays it alone.  Here
// all decoder handles of one (code, algorithm, precision) share a MAILBOX:
//   * a small arena of slots in MAPPED pinned host memory: per slot the frame's float64 LLRs (the CPU copies them in, the
//     GPU reads them over PCIe), and its packed decisions / iteration count / flag / completion word;
//   * a caller copies its LLRs into a free slot, queues the slot, and whoever finds the launch lock free becomes the
//     LEADER: it takes everything queued so far (frames of other threads included -- that is the coalescing), and issues
//     ONE launch for them on one of a few streams; leadership ends with the launch, not with the completion, so the next
//     batch can be issued while this one runs;
//   * the kernel (latency instance: one frame per CTA spread over NP = 4 parts, static grid, no work counter) reads the
//     float64 row straight from the slot, decodes, writes the decisions back into the slot and publishes the completion
//     word behind a system-scope fence; every caller spins on ITS slot's word -- no stream synchronisation, no copy engine.
// Results are those of the batched path bit for bit (same kernels' arithmetic, same float64 -> float32 rounding).
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>
#include <immintrin.h>

namespace ldpc {

#include "resident_common.cuh"

static constexpr int MB_SLOTS = 64;        // <= MAILBOX_MAX_BATCH and <= 64 (one bit each in the free mask)
static constexpr int MB_STREAMS = 4;
static_assert(MB_SLOTS <= MAILBOX_MAX_BATCH && MB_SLOTS <= 64, "slot map / free mask size");

struct MbPending { int slot, max_iters, take; };

struct Mailbox {
    const ldpc_b200_code *code = nullptr;
    int device = 0, algo = 0, precision = 0, engine = 0, N = 0, W = 0;
    double *h_llr = nullptr;               // [MB_SLOTS][N]   mapped
    uint32_t *h_bits = nullptr;            // [MB_SLOTS][W]   mapped
    int32_t *h_iters = nullptr;            // [MB_SLOTS]
    uint8_t *h_conv = nullptr;             // [MB_SLOTS]
    unsigned int *h_done = nullptr;        // [MB_SLOTS] completion word: 0 = pending, 0xffffffff = launch failed, else the batch number
    std::atomic<uint64_t> free_mask{~0ull};
    std::mutex q_mu;
    std::vector<MbPending> pending;
    std::atomic<int> n_pending{0};         // = pending.size(), readable without the lock (waiters lead only when something is queued)
    std::atomic<int> preparing{0};         // callers that own a slot but have not queued it yet (the leader's bounded wait)
    std::mutex launch_mu;
    cudaStream_t streams[MB_STREAMS] = {};
    int next_stream = 0;
    unsigned int seq = 0;
    ldpc_b200_decoder ldec;                // launch context only (code / algo / precision / engine); owns no buffers
    std::string launch_error;
    int refs = 0;
    std::atomic<uint64_t> n_launches{0}, n_frames{0};
    // LDPC_B200_MB_PROFILE=1: where the time of a call goes (ns, summed over calls; printed when the mailbox is freed)
    std::atomic<uint64_t> t_copy{0}, t_queue_launch{0}, t_wait{0}, t_unpack{0}, t_calls{0};
};

static bool mb_profile()
{
    static const bool on = [] { const char *e = getenv("LDPC_B200_MB_PROFILE"); return e && e[0] == '1'; }();
    return on;
}

static std::mutex g_mb_mu;
static std::map<std::tuple<const ldpc_b200_code *, int, int, int>, Mailbox *> g_mailboxes;

static void mailbox_free(Mailbox *m)
{
    if (mb_profile() && m->t_calls.load()) {
        const double n = (double)m->t_calls.load();
        fprintf(stderr, "mailbox profile (N=%d, %llu calls, %llu launches): copy-in %.2f us, queue+launch %.2f us, wait %.2f us, unpack %.2f us per call\n",
                m->N, (unsigned long long)m->t_calls.load(), (unsigned long long)m->n_launches.load(), m->t_copy.load() / n / 1e3,
                m->t_queue_launch.load() / n / 1e3, m->t_wait.load() / n / 1e3, m->t_unpack.load() / n / 1e3);
    }
    cudaSetDevice(m->device);
    for (int i = 0; i < MB_STREAMS; i++) if (m->streams[i]) { cudaStreamSynchronize(m->streams[i]); cudaStreamDestroy(m->streams[i]); }
    if (m->h_llr) cudaFreeHost(m->h_llr);
    if (m->h_bits) cudaFreeHost(m->h_bits);
    if (m->h_iters) cudaFreeHost(m->h_iters);
    if (m->h_conv) cudaFreeHost(m->h_conv);
    if (m->h_done) cudaFreeHost(m->h_done);
    delete m;
}

static int mailbox_create(ldpc_b200_decoder *d, int engine, Mailbox **out)
{
    Mailbox *m = new (std::nothrow) Mailbox();
    if (!m) return LDPC_B200_ENOMEM;
    m->code = d->code; m->device = d->device; m->algo = d->algo; m->precision = d->precision; m->engine = engine;
    m->N = d->code->N; m->W = words_per_frame(m->N);
    m->ldec.code = d->code; m->ldec.algo = d->algo; m->ldec.precision = d->precision; m->ldec.engine = engine;
    m->ldec.device = d->device; m->ldec.max_batch = MB_SLOTS;
    cudaError_t e = cudaSetDevice(m->device);
    // write-combined slots were measured against ordinary (cached) pinned memory: the same for one caller, 3-10 % slower from 16
    // threads up (16 threads, 2.0 dB: 196-197 k frames/s against 203-218 k) -- the default is cached, LDPC_B200_MB_WC=1 switches
    static const bool wc = [] { const char *e = getenv("LDPC_B200_MB_WC"); return e && e[0] == '1'; }();
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&m->h_llr, (size_t)MB_SLOTS * m->N * sizeof(double), cudaHostAllocMapped | (wc ? cudaHostAllocWriteCombined : 0));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&m->h_bits, (size_t)MB_SLOTS * m->W * sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&m->h_iters, MB_SLOTS * sizeof(int32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&m->h_conv, MB_SLOTS, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&m->h_done, MB_SLOTS * sizeof(unsigned int), cudaHostAllocMapped);
    for (int i = 0; i < MB_STREAMS && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) { mailbox_free(m); return cuda_fail(e, "mailbox: mapped host memory / streams", __FILE__, __LINE__); }
    memset(m->h_done, 0, MB_SLOTS * sizeof(unsigned int));
    m->ldec.stream = m->streams[0];
    *out = m;
    return 0;
}

// The mailbox of this decoder's (code, algorithm, precision, engine), created on first use and shared by all handles.
static int mailbox_get(ldpc_b200_decoder *d, int engine, Mailbox **out)
{
    if (d->mbox && ((Mailbox *)d->mbox)->engine == engine) { *out = (Mailbox *)d->mbox; return 0; }
    std::lock_guard<std::mutex> lk(g_mb_mu);
    if (d->mbox) {                                        // the engine of the handle was switched: move to the other mailbox
        Mailbox *old = (Mailbox *)d->mbox;
        if (--old->refs == 0) { g_mailboxes.erase(std::make_tuple(old->code, old->algo, old->precision, old->engine)); mailbox_free(old); }
        d->mbox = nullptr;
    }
    auto key = std::make_tuple(d->code, d->algo, d->precision, engine);
    auto it = g_mailboxes.find(key);
    Mailbox *m;
    if (it == g_mailboxes.end()) {
        int rc = mailbox_create(d, engine, &m);
        if (rc) return rc;
        g_mailboxes[key] = m;
    } else m = it->second;
    m->refs++;
    d->mbox = m;
    *out = m;
    return 0;
}

void mailbox_release(ldpc_b200_decoder *d)
{
    if (!d->mbox) return;
    std::lock_guard<std::mutex> lk(g_mb_mu);
    Mailbox *m = (Mailbox *)d->mbox;
    d->mbox = nullptr;
    if (--m->refs == 0) {
        g_mailboxes.erase(std::make_tuple(m->code, m->algo, m->precision, m->engine));
        mailbox_free(m);
    }
}

// n slots at once, all or nothing (one compare-and-swap): a caller never holds some slots while it waits for others, so
// callers with several frames each cannot deadlock on a full arena.
static void acquire_slots(Mailbox *m, int n, int *slots)
{
    for (int spin = 0;; spin++) {
        uint64_t mask = m->free_mask.load(std::memory_order_acquire);
        while (__builtin_popcountll(mask) >= n) {
            uint64_t take = 0, rest = mask;
            for (int k = 0; k < n; k++) { const int s = __builtin_ctzll(rest); slots[k] = s; take |= 1ull << s; rest &= rest - 1; }
            if (m->free_mask.compare_exchange_weak(mask, mask & ~take, std::memory_order_acq_rel)) return;
        }
        if (spin < 64) _mm_pause(); else std::this_thread::yield();       // arena in flight: more frames than slots
    }
}

// Leader: one launch per (max_iters, take) group of everything queued.  Called with launch_mu held.
static void launch_pending(Mailbox *m)
{
    // bounded wait (<= 100 pause instructions, ~4 us): callers that already own a slot are about to queue it -- take them along.
    // Measured on a 16-vCPU B200 host, jpl.4096.4.5 at 2.0 dB, 16 threads: no wait 129 k frames/s (1.6 frames per launch: the
    // launches themselves, ~5.5 us each under the lock, become the limit), 24: 196 k, 100: 214 k, 400: 185 k.
    static const int wait_pauses = [] { const char *e = getenv("LDPC_B200_MB_WAIT"); return e ? atoi(e) : 100; }();
    for (int spin = 0; spin < wait_pauses && m->preparing.load(std::memory_order_acquire) > 0; spin++) _mm_pause();
    std::vector<MbPending> batch;
    { std::lock_guard<std::mutex> lk(m->q_mu); batch.swap(m->pending); m->n_pending.store(0, std::memory_order_release); }
    if (batch.empty()) return;
    cudaSetDevice(m->device);
    std::vector<char> used(batch.size(), 0);
    for (size_t i = 0; i < batch.size(); i++) {
        if (used[i]) continue;
        ResidentParams p = {};
        p.llr = nullptr; p.llr_bulk = 0;
        p.llr64 = m->h_llr; p.llr64_stride = m->N; p.llr64_take = batch[i].take;        // (UVA: mapped host pointers are device pointers)
        p.bits = m->h_bits; p.iters = m->h_iters; p.conv = m->h_conv; p.counters = nullptr;
        p.max_iters = batch[i].max_iters; p.early_exit = 1; p.k_info = 0; p.n_tx = m->N;
        p.use_map = 1; p.done = m->h_done;
        unsigned int n = 0;
        for (size_t j = i; j < batch.size(); j++)
            if (!used[j] && batch[j].max_iters == batch[i].max_iters && batch[j].take == batch[i].take) { used[j] = 1; p.slot_map[n++] = (unsigned short)batch[j].slot; }
        p.n_frames = n;
        if (++m->seq == 0 || m->seq == 0xffffffffu) m->seq = 1;
        p.done_seq = m->seq;
        cudaStream_t st = m->streams[m->next_stream];
        m->next_stream = (m->next_stream + 1) % MB_STREAMS;
        int rc = resident_latency_supports(&m->ldec) ? resident_latency_launch(&m->ldec, p, st) : resident_static_launch(&m->ldec, p, st);
        if (rc) {
            m->launch_error = ldpc_b200_last_error();
            for (unsigned int q = 0; q < n; q++) { __atomic_store_n(&m->h_done[p.slot_map[q]], 0xffffffffu, __ATOMIC_RELEASE); }
        }
        m->n_launches.fetch_add(1, std::memory_order_relaxed);
        m->n_frames.fetch_add(n, std::memory_order_relaxed);
    }
}

bool mailbox_enabled()
{
    static const bool on = [] { const char *e = getenv("LDPC_B200_MAILBOX"); return !(e && e[0] == '0'); }();
    return on;
}

// n_frames (<= a few) float64 rows through the mailbox.  llr: rows of `take` values at a stride of `stride` doubles.
// Outputs as in decode_host: bits_bytes rows of out_take decisions (0 = N), or packed words; iters / converged may be NULL.
int mailbox_decode(ldpc_b200_decoder *d, int engine, int max_iters, int n_frames, const double *llr, int stride, int take,
                   uint8_t *bits_bytes, int out_take, uint32_t *bits_packed, int32_t *iters, uint8_t *converged)
{
    Mailbox *m;
    int rc = mailbox_get(d, engine, &m);
    if (rc) return rc;
    const int N = m->N, W = m->W;
    if (take <= 0 || take > N) take = N;
    if (stride <= 0) stride = N;
    if (out_take <= 0 || out_take > N) out_take = N;
    int slots[MB_SLOTS];
    if (n_frames > MB_SLOTS / 4) { set_error("mailbox_decode: batch of %d frames belongs on the batched path", n_frames); return LDPC_B200_EINVAL; }
    const bool prof = mb_profile();
    auto now_ns = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const uint64_t p0 = prof ? now_ns() : 0;
    acquire_slots(m, n_frames, slots);
    m->preparing.fetch_add(1, std::memory_order_acq_rel);
    for (int f = 0; f < n_frames; f++) {
        const int s = slots[f];
        memcpy(m->h_llr + (size_t)s * N, llr + (size_t)f * stride, (size_t)take * sizeof(double));
        __atomic_store_n(&m->h_done[s], 0u, __ATOMIC_RELAXED);
    }
    _mm_sfence();                                       // (write-combined slots: the stores are globally visible before the slot is queued)
    {
        std::lock_guard<std::mutex> lk(m->q_mu);
        for (int f = 0; f < n_frames; f++) m->pending.push_back({slots[f], max_iters, take});
        m->n_pending.store((int)m->pending.size(), std::memory_order_release);
    }
    m->preparing.fetch_sub(1, std::memory_order_acq_rel);
    const uint64_t p1 = prof ? now_ns() : 0;
    uint64_t p2 = 0, p3 = 0;                              // first launch attempt done / all completion words seen
    if (prof && m->n_pending.load(std::memory_order_acquire) > 0 && m->launch_mu.try_lock()) { launch_pending(m); m->launch_mu.unlock(); }
    if (prof) p2 = now_ns();

    rc = 0;
    // spin briefly, then yield on every probe: with more callers than cores a spinning waiter would otherwise keep the
    // thread that has to copy / launch off the CPU (32 threads on 16 cores: 30-57 k frames/s when spinning 4096 probes,
    // 318 k when yielding after 64; with no more threads than cores the yield returns at once and costs nothing measurable)
    static const unsigned long spin_limit = [] { const char *e = getenv("LDPC_B200_MB_SPIN"); return (unsigned long)(e ? atol(e) : 64); }();
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < n_frames; f++) {
        const int s = slots[f];
        unsigned int v;
        for (unsigned long spin = 0; (v = __atomic_load_n(&m->h_done[s], __ATOMIC_ACQUIRE)) == 0; spin++) {
            // not done: if nobody is launching, lead (our slot, or later arrivals, may still be queued)
            if (m->n_pending.load(std::memory_order_acquire) > 0 && m->launch_mu.try_lock()) { launch_pending(m); m->launch_mu.unlock(); }
            if (spin < spin_limit) _mm_pause();
            else {
                std::this_thread::yield();
                if ((spin & 4095) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                    cudaError_t e = cudaStreamQuery(m->streams[0]);
                    set_error("mailbox: no completion after 20 s (stream state: %s)", cudaGetErrorString(e));
                    rc = LDPC_B200_ECUDA;
                    break;
                }
            }
        }
        if (rc) break;
        if (v == 0xffffffffu) { set_error("mailbox launch failed: %s", m->launch_error.c_str()); rc = LDPC_B200_ECUDA; break; }
        if (prof && f == n_frames - 1) p3 = now_ns();
        const uint32_t *pk = m->h_bits + (size_t)s * W;
        if (bits_bytes) {                              // packed word -> one byte per decision, 8 at a time
            uint8_t *o = bits_bytes + (size_t)f * out_take;
            const uint8_t *pb = reinterpret_cast<const uint8_t *>(pk);
            int j = 0;
            for (; j + 8 <= out_take; j += 8) {
                const uint64_t x = ((uint64_t)pb[j >> 3] * 0x0101010101010101ull) & 0x8040201008040201ull;   // byte i keeps bit i
                const uint64_t y = ((x + 0x7f7f7f7f7f7f7f7full) >> 7) & 0x0101010101010101ull;              // non-zero byte -> 1
                memcpy(o + j, &y, 8);
            }
            for (; j < out_take; j++) o[j] = (pk[j >> 5] >> (j & 31)) & 1u;
        }
        if (bits_packed) memcpy(bits_packed + (size_t)f * W, pk, (size_t)W * sizeof(uint32_t));
        if (iters) iters[f] = m->h_iters[s];
        if (converged) converged[f] = m->h_conv[s];
    }
    // slots go back only when the kernel can no longer touch them: completed, or never launched successfully
    uint64_t give = 0;
    for (int f = 0; f < n_frames; f++) if (rc == 0 || __atomic_load_n(&m->h_done[slots[f]], __ATOMIC_ACQUIRE) != 0) give |= 1ull << slots[f];
    m->free_mask.fetch_or(give, std::memory_order_acq_rel);
    if (prof && !rc) {
        const uint64_t p4 = now_ns();
        m->t_copy += p1 - p0; m->t_queue_launch += p2 - p1; m->t_wait += p3 - p2; m->t_unpack += p4 - p3; m->t_calls += 1;
    }
    return rc;
}

void mailbox_stats(const ldpc_b200_decoder *d, uint64_t *launches, uint64_t *frames)
{
    const Mailbox *m = (const Mailbox *)d->mbox;
    if (launches) *launches = m ? m->n_launches.load() : 0;
    if (frames) *frames = m ? m->n_frames.load() : 0;
}

}  // namespace ldpc
