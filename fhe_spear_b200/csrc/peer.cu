// peer.cu -- combining the shard accumulators of a giant-step-sharded mat-vec over NVLink peer memory.
//
// every rank ends its share of the giant steps with an accumulator R_r in basis Q_l*P;
// the mat-vec needs  R = sum_r R_r  (mod q per limb).  NCCL's sum is not mod q, and going through it costs two
// host hand-offs between the engine's stream and the communicator's.  Here the exchange is part of the engine:
// every rank owns a *window* (cudaMalloc memory exported with cudaIpc, mapped by the other ranks of its group),
// and one kernel per rank does reduce-scatter + modular reduction + all-gather in a single pass:
//
//     rank r owns slice r of the accumulator;  for every element of its slice it loads the R contributions
//     straight from the R windows (R-1 of them over NVLink), adds them, reduces mod q_limb (one Barrett step:
//     <= 8 residues < 2^60 cannot wrap 64 bits) and stores the result back into slice r of every window.
//
// Ordering is carried by epoch flags in the windows (st.release.sys / ld.acquire.sys), posted and awaited by
// one-warp kernels on the engine's stream -- no host synchronisation, no second communicator stream:
//
//     copy R_r -> window | post ready[e] to all peers, wait ready[e] from all | k_peer_reduce (last CTA posts
//     done[e] to all peers) | wait done[e] from all | copy window -> R
//
// Reuse is safe without double buffering: a peer reads my window only between my ready[e] and its done[e], and
// writes only its own slice of it in that interval; I touch the window again only after every done[e] arrived.
// A peer that never arrives does not hang the GPU: waits give up after SPIN_TIMEOUT_NS and raise the window's
// status word (host-mapped), which makes this and every later call on the window fail.
#include <cstring>
#include <exception>
#include <string>

#include "../../include/spear_b200.h"
#include "engine.h"
#include "ops.h"

namespace {

constexpr int MAX_PEERS = 8, MAX_SLOTS = 8, SLOT_FLAG_WORDS = 32, DATA_OFFSET_WORDS = 512;
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

struct PeerWindow : CtxRef {
    int rank = 0, world = 1, slots = 0;
    size_t slot_words = 0;
    u64* base = nullptr;               // [DATA_OFFSET_WORDS flags][slots][slot_words]
    u64* peer[MAX_PEERS] = {};         // mapped windows of the group (peer[rank] == base)
    bool connected = false;
    int mode = 0;                      // 0: unused yet, 1: accumulator all-reduce, 2: two-phase mat-vec exchange (never mixed)
    u64 epoch[MAX_SLOTS] = {};
    int* status = nullptr;             // pinned host word, written by the waiting kernels on time-out
    int* d_status = nullptr;           // its device address (zero-copy: only the one-warp sync kernels touch it)
    int* d_failed = nullptr;           // the same flag in device memory, read by the reduce / collect kernels
    ~PeerWindow() {
        if (!ctx) return;
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        for (int r = 0; r < world; r++)
            if (r != rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
        if (base) cudaFree(base);
        if (d_failed) cudaFree(d_failed);
        if (status) cudaFreeHost(status);
    }
};

struct PeerPtrs {
    u64* w[MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 ld_sys_v2(const u64* p) {   // coherent at system scope: never a stale L1 line
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_v2(u64* p, ulonglong2 v) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// flag word of (slot, kind, writer) inside a window: kind 0 = ready, 1 = done
__host__ __device__ __forceinline__ size_t flag_index(int slot, int kind, int writer) {
    return (size_t)slot * SLOT_FLAG_WORDS + (size_t)kind * MAX_PEERS + writer;
}

// lane r < world:  (post != 0) tell peer r that `me` reached `epoch` of `kind`;  then (wait != 0) wait until peer r told
// us the same
__global__ void k_peer_sync(PeerPtrs pp, int world, int me, int slot, int kind, int post, u64 epoch, int* status,
                            int* failed, int wait = 1) {
    const int r = threadIdx.x;
    if (r >= world) return;
    if (post) {
        __threadfence_system();
        st_release_sys(pp.w[r] + flag_index(slot, kind, me), epoch);
    }
    if (!wait) return;
    const u64* mine = pp.w[me] + flag_index(slot, kind, r);
    const unsigned long long deadline = globaltimer_ns() + SPIN_TIMEOUT_NS;
    while (ld_acquire_sys(mine) < epoch) {
        if (globaltimer_ns() > deadline) {
            *(volatile int*)status = 1 + r;   // host-visible (zero-copy)
            *(volatile int*)failed = 1 + r;   // device copy for the kernels queued behind this one
            __threadfence_system();
            return;
        }
        __nanosleep(100);
    }
}

// slice [lo, hi) (in pairs of words) of the accumulator: sum over the R windows, mod q, written back to every window
template <int R>
__global__ void __launch_bounds__(256) k_peer_reduce(PeerPtrs pp, size_t data_off, size_t lo, size_t hi, int rows,
                                                     int logn, RowMap rm, ModTab mt, int me, int slot, u64 epoch,
                                                     const int* failed) {
    if (*(volatile const int*)failed) return;   // a peer never arrived: its window holds stale data, touch nothing
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < hi; v += stride) {
        const size_t e = data_off + 2 * v;
        ulonglong2 x[R];
#pragma unroll
        for (int r = 0; r < R; r++) x[r] = ld_sys_v2(pp.w[r] + e);
        ulonglong2 s = x[0];
#pragma unroll
        for (int r = 1; r < R; r++) s.x += x[r].x, s.y += x[r].y;
        const int limb = rm.limb((int)(((2 * v) >> logn) % rows));
        const u64 q = mt.q[limb], r1 = mt.ratio1[limb];
        s.x = barrett64(s.x, q, r1), s.y = barrett64(s.y, q, r1);
#pragma unroll
        for (int r = 0; r < R; r++) st_sys_v2(pp.w[r] + e, s);
    }
    // the last CTA to finish tells every peer that slice `me` is complete
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        u64* counter = pp.w[me] + (size_t)slot * SLOT_FLAG_WORDS + 2 * MAX_PEERS;
        const unsigned long long prev = atomicAdd((unsigned long long*)counter, 1ull);
        if (prev == gridDim.x - 1) {
            atomicExch((unsigned long long*)counter, 0ull);
            __threadfence_system();
#pragma unroll
            for (int r = 0; r < R; r++) st_release_sys(pp.w[r] + flag_index(slot, 1, me), epoch);
        }
    }
}

// window -> accumulator, or -- when a wait timed out -- an accumulator of all-ones words (no valid residue: q < 2^61),
// so that a failed exchange can never pass for a ciphertext downstream
__global__ void k_peer_collect(const u64* __restrict__ win, u64* __restrict__ acc, size_t words, const int* failed_flag) {
    const bool failed = *(volatile const int*)failed_flag != 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < words; e += (size_t)gridDim.x * blockDim.x)
        acc[e] = failed ? ~0ull : win[e];
}

inline PeerWindow* W_(spear_peer_window* w) { return reinterpret_cast<PeerWindow*>(w); }

// R <- all-ones words when a wait of this window timed out (no valid residue: q < 2^61)
__global__ void k_peer_poison(u64* __restrict__ R, size_t words, const int* failed_flag) {
    if (*(volatile const int*)failed_flag == 0) return;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < words; e += (size_t)gridDim.x * blockDim.x) R[e] = ~0ull;
}

}  // namespace

// ---- two-phase mat-vec: the exchange between the row-split phase and the giant-group-split phase -----------------------
// Flags of a slot: kind 0 = "my phase-1 stores into your slot are complete" (epoch e), kind 1 = "I have consumed my slot
// of epoch e".  A rank may write into a peer's slot for epoch e only after that peer consumed epoch e - 1.
namespace peer {
static PeerPtrs ptrs_of(const PeerWindow* w) {
    PeerPtrs pp;
    for (int r = 0; r < MAX_PEERS; r++) pp.w[r] = w->peer[r < w->world ? r : w->rank];
    return pp;
}
void window_geometry(const spear_peer_window* win, int* rank, int* world) {
    const PeerWindow* w = reinterpret_cast<const PeerWindow*>(win);
    REQUIRE(w, "two-phase mat-vec: null window");
    *rank = w->rank, *world = w->world;
}
SplitView split_begin(const Ctx* c, spear_peer_window* win, int slot, size_t need_words, cudaStream_t s) {
    PeerWindow* w = W_(win);
    REQUIRE(w && w->connected && w->ctx == c, "two-phase mat-vec: window not connected");
    REQUIRE(w->mode == 0 || w->mode == 2, "two-phase mat-vec: this window serves accumulator all-reduces");
    REQUIRE(slot >= 0 && slot < w->slots, "two-phase mat-vec: slot %d of %d", slot, w->slots);
    REQUIRE(need_words <= w->slot_words, "two-phase mat-vec: the window slot holds %zu words, %zu needed", w->slot_words, need_words);
    REQUIRE(*(volatile int*)w->status == 0, "two-phase mat-vec: peer %d never arrived (CUDA peer exchange timed out)",
            *(volatile int*)w->status - 1);
    w->mode = 2;
    const u64 epoch = ++w->epoch[slot];
    ProfScope ps(c, PROF_PEER_WAIT, s);
    if (w->world > 1)
        LAUNCH(k_peer_sync, 1, 32, 0, s)(ptrs_of(w), w->world, w->rank, slot, 1, 0, epoch - 1, w->d_status, w->d_failed, 1);
    SplitView v;
    v.rank = w->rank, v.world = w->world;
    const size_t data_off = DATA_OFFSET_WORDS + (size_t)slot * w->slot_words;
    for (int r = 0; r < MAX_PEERS; r++) v.base[r] = w->peer[r < w->world ? r : w->rank] + data_off;
    return v;
}
void split_exchange(spear_peer_window* win, int slot, cudaStream_t s) {
    PeerWindow* w = W_(win);
    ProfScope ps(w->ctx, PROF_PEER_WAIT, s);
    if (w->world > 1)
        LAUNCH(k_peer_sync, 1, 32, 0, s)(ptrs_of(w), w->world, w->rank, slot, 0, 1, w->epoch[slot], w->d_status, w->d_failed, 1);
}
void split_post(spear_peer_window* win, int slot, cudaStream_t s) {
    PeerWindow* w = W_(win);
    if (w->world > 1)
        LAUNCH(k_peer_sync, 1, 32, 0, s)(ptrs_of(w), w->world, w->rank, slot, 0, 1, w->epoch[slot], w->d_status, w->d_failed, 0);
}
void split_wait(spear_peer_window* win, int slot, cudaStream_t s) {
    PeerWindow* w = W_(win);
    ProfScope ps(w->ctx, PROF_PEER_WAIT, s);
    if (w->world > 1)
        LAUNCH(k_peer_sync, 1, 32, 0, s)(ptrs_of(w), w->world, w->rank, slot, 0, 0, w->epoch[slot], w->d_status, w->d_failed, 1);
}
void split_release(spear_peer_window* win, int slot, u64* R, size_t words, cudaStream_t s) {
    PeerWindow* w = W_(win);
    if (w->world > 1) {
        LAUNCH(k_peer_sync, 1, 32, 0, s)(ptrs_of(w), w->world, w->rank, slot, 1, 1, w->epoch[slot], w->d_status, w->d_failed, 0);
        LAUNCH(k_peer_poison, (int)std::min<size_t>((words + 255) / 256, (size_t)w->ctx->sm_count * 4), 256, 0, s)(R, words, w->d_failed);
    }
    CUDA_CHECK(cudaGetLastError());
}
}  // namespace peer

void spear_set_last_error(const char* msg);   // api.cu

#define PEER_BEGIN try {
#define PEER_END                                    \
    }                                               \
    catch (const spear_error& e) {                  \
        spear_set_last_error(e.msg);                \
        return e.code ? e.code : SPEAR_ERR_INVALID; \
    }                                               \
    catch (const std::exception& e) {               \
        spear_set_last_error(e.what());             \
        return SPEAR_ERR_INVALID;                   \
    }                                               \
    return SPEAR_OK;

extern "C" {

int spear_peer_window_create(spear_context* ctx, int rank, int world, uint64_t slot_bytes, int slots,
                             uint8_t* handle_out, spear_peer_window** out) {
    PEER_BEGIN
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, "peer window: rank %d of %d", rank, world);
    REQUIRE(slots >= 1 && slots <= MAX_SLOTS && slot_bytes > 0, "peer window: 1..%d slots", MAX_SLOTS);
    static_assert(sizeof(cudaIpcMemHandle_t) == SPEAR_IPC_HANDLE_BYTES, "IPC handle size");
    CUDA_CHECK(cudaSetDevice(c->device));
    std::unique_ptr<PeerWindow> w(new PeerWindow);
    w->bind(c), w->rank = rank, w->world = world, w->slots = slots;
    w->slot_words = ((slot_bytes + 7) / 8 + 31) & ~(size_t)31;
    const size_t words = DATA_OFFSET_WORDS + (size_t)slots * w->slot_words;
    CUDA_CHECK(cudaMalloc(&w->base, words * sizeof(u64)));   // cudaIpc needs cudaMalloc memory, not the async pool
    CUDA_CHECK(cudaMemset(w->base, 0, DATA_OFFSET_WORDS * sizeof(u64)));
    CUDA_CHECK(cudaHostAlloc(&w->status, sizeof(int), cudaHostAllocMapped));
    *w->status = 0;
    CUDA_CHECK(cudaHostGetDevicePointer(&w->d_status, w->status, 0));
    CUDA_CHECK(cudaMalloc(&w->d_failed, sizeof(int)));
    CUDA_CHECK(cudaMemset(w->d_failed, 0, sizeof(int)));
    w->peer[rank] = w->base;
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof(h));
    if (world > 1) CUDA_CHECK(cudaIpcGetMemHandle(&h, w->base));
    std::memcpy(handle_out, &h, sizeof(h));
    CUDA_CHECK(cudaDeviceSynchronize());
    w->connected = world == 1;
    *out = reinterpret_cast<spear_peer_window*>(w.release());
    PEER_END
}

int spear_peer_window_connect(spear_context* ctx, spear_peer_window* win, const uint8_t* handles) {
    PEER_BEGIN
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    PeerWindow* w = W_(win);
    REQUIRE(w && !w->connected, "peer window: already connected");
    CUDA_CHECK(cudaSetDevice(c->device));
    for (int r = 0; r < w->world; r++) {
        if (r == w->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * SPEAR_IPC_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        w->peer[r] = static_cast<u64*>(p);
    }
    w->connected = true;
    PEER_END
}

int spear_peer_allreduce(spear_context* ctx, spear_peer_window* win, int slot, spear_obj* acc_) {
    PEER_BEGIN
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    PeerWindow* w = W_(win);
    Obj* acc = reinterpret_cast<Obj*>(acc_);
    REQUIRE(w && w->connected, "peer all-reduce: window not connected");
    REQUIRE(w->mode == 0 || w->mode == 1, "peer all-reduce: this window serves the two-phase mat-vec exchange");
    w->mode = 1;
    REQUIRE(slot >= 0 && slot < w->slots, "peer all-reduce: slot %d of %d", slot, w->slots);
    REQUIRE(acc && acc->n == c->N && acc->words() % 2 == 0, "peer all-reduce: bad operand");
    REQUIRE(acc->words() <= w->slot_words, "peer all-reduce: operand larger than the window slot");
    REQUIRE(*(volatile int*)w->status == 0, "peer all-reduce: peer %d never arrived (CUDA peer exchange timed out)",
            *(volatile int*)w->status - 1);
    CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    if (w->world == 1) {
        ops::reduce_inplace(c, acc->d, acc->size, acc->rows(), RowMap{acc->rows(), acc->l, c->L, 0}, s);
        return SPEAR_OK;
    }
    const u64 epoch = ++w->epoch[slot];
    const size_t data_off = DATA_OFFSET_WORDS + (size_t)slot * w->slot_words, bytes = acc->words() * sizeof(u64);
    PeerPtrs pp;
    for (int r = 0; r < MAX_PEERS; r++) pp.w[r] = w->peer[r < w->world ? r : w->rank];
    {
        ProfScope ps(c, PROF_PEER_REDUCE, s);
        CUDA_CHECK(cudaMemcpyAsync(w->base + data_off, acc->d, bytes, cudaMemcpyDeviceToDevice, s));
    }
    {
        ProfScope ps(c, PROF_PEER_WAIT, s);
        LAUNCH(k_peer_sync, 1, 32, 0, s)(pp, w->world, w->rank, slot, 0, 1, epoch, w->d_status, w->d_failed, 1);
    }
    const size_t pairs = acc->words() / 2, per = (pairs + w->world - 1) / w->world;
    const size_t lo = std::min(pairs, per * w->rank), hi = std::min(pairs, lo + per);
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((hi - lo + 255) / 256, (size_t)c->sm_count * 8));
    const RowMap rm{acc->rows(), acc->l, c->L, 0};
    auto go = [&](auto kern) {
        LAUNCH(kern, grid, 256, 0, s)(pp, data_off, lo, hi, acc->rows(), c->logn, rm, c->modtab(), w->rank, slot, epoch,
                                      w->d_failed);
    };
    {
        ProfScope ps(c, PROF_PEER_REDUCE, s);
        switch (w->world) {
            case 2: go(k_peer_reduce<2>); break;
            case 3: go(k_peer_reduce<3>); break;
            case 4: go(k_peer_reduce<4>); break;
            case 5: go(k_peer_reduce<5>); break;
            case 6: go(k_peer_reduce<6>); break;
            case 7: go(k_peer_reduce<7>); break;
            default: go(k_peer_reduce<8>); break;
        }
    }
    {
        ProfScope ps(c, PROF_PEER_WAIT, s);
        LAUNCH(k_peer_sync, 1, 32, 0, s)(pp, w->world, w->rank, slot, 1, 0, epoch, w->d_status, w->d_failed, 1);
    }
    ProfScope ps(c, PROF_PEER_REDUCE, s);
    LAUNCH(k_peer_collect, (int)std::min<size_t>((acc->words() + 255) / 256, (size_t)c->sm_count * 8), 256, 0, s)(
        w->base + data_off, acc->d, acc->words(), w->d_failed);
    CUDA_CHECK(cudaGetLastError());
    PEER_END
}

// Test hook (one GPU, one process): the reduce kernels of all `world` ranks run one after the other over `world`
// local windows holding accs[0..world), without the epoch waits -- kernels that wait on one another must not share a
// GPU as separate launches.  Every accs[r] ends as the sum of all of them mod q, exactly as after a real exchange.
int spear_peer_selftest(spear_context* ctx, spear_obj* const* accs_, int world) {
    PEER_BEGIN
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    REQUIRE(world >= 2 && world <= MAX_PEERS, "peer selftest: 2..%d ranks", MAX_PEERS);
    CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    Obj* const* accs = reinterpret_cast<Obj* const*>(accs_);
    const size_t words = accs[0]->words();
    for (int r = 0; r < world; r++)
        REQUIRE(accs[r] && accs[r]->words() == words && accs[r]->n == c->N && words % 2 == 0, "peer selftest: bad operand");
    PeerPtrs pp;
    u64* win[MAX_PEERS] = {};
    int* d_status = (int*)c->alloc(8, s);
    CUDA_CHECK(cudaMemsetAsync(d_status, 0, 8, s));
    for (int r = 0; r < world; r++) {
        win[r] = c->alloc(DATA_OFFSET_WORDS + words, s);
        CUDA_CHECK(cudaMemsetAsync(win[r], 0, DATA_OFFSET_WORDS * sizeof(u64), s));
        CUDA_CHECK(cudaMemcpyAsync(win[r] + DATA_OFFSET_WORDS, accs[r]->d, words * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    }
    for (int r = 0; r < MAX_PEERS; r++) pp.w[r] = win[r < world ? r : 0];
    const Obj* a0 = accs[0];
    const RowMap rm{a0->rows(), a0->l, c->L, 0};
    const size_t pairs = words / 2, per = (pairs + world - 1) / world;
    for (int me = 0; me < world; me++) {
        const size_t lo = std::min(pairs, per * me), hi = std::min(pairs, lo + per);
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((hi - lo + 255) / 256, (size_t)c->sm_count * 8));
        auto go = [&](auto kern) {
            LAUNCH(kern, grid, 256, 0, s)(pp, (size_t)DATA_OFFSET_WORDS, lo, hi, a0->rows(), c->logn, rm, c->modtab(), me, 0,
                                          (u64)1, d_status);
        };
        switch (world) {
            case 2: go(k_peer_reduce<2>); break;
            case 3: go(k_peer_reduce<3>); break;
            case 4: go(k_peer_reduce<4>); break;
            case 5: go(k_peer_reduce<5>); break;
            case 6: go(k_peer_reduce<6>); break;
            case 7: go(k_peer_reduce<7>); break;
            default: go(k_peer_reduce<8>); break;
        }
    }
    for (int r = 0; r < world; r++) {
        LAUNCH(k_peer_collect, (int)std::min<size_t>((words + 255) / 256, (size_t)c->sm_count * 8), 256, 0, s)(
            win[r] + DATA_OFFSET_WORDS, accs[r]->d, words, d_status);
        c->free(win[r], s);
    }
    c->free(d_status, s);
    CUDA_CHECK(cudaGetLastError());
    PEER_END
}

int spear_peer_window_status(const spear_peer_window* win) {
    const PeerWindow* w = reinterpret_cast<const PeerWindow*>(win);
    return w && w->status ? *(volatile int*)w->status : -1;
}

void spear_peer_window_destroy(spear_peer_window* win) { delete W_(win); }

}  // extern "C"
