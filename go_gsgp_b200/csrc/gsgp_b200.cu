// gsgp_b200.cu -- context + C-ABI (include/gsgp_b200.h) over the sm_100a kernels.
//
// The context owns: the dataset shard (variable-major X, targets y), the double-buffered semantic
// store (init_tables main_cpu.go:1257-1284; update_tables :1191-1197 is a buffer flip), the fitness
// tables, the generation plan + random trees + their postfix programs, the random-tree semantic
// workspace R, the per-tile error partials, and one stream on which everything is issued in order.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: libnccl is dlopen()ed when world_size > 1

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "../../include/gsgp_host.h"
#include "trees.hpp"

using namespace gsgp;

namespace {

thread_local std::string g_create_error;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err)
    {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
        AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllGather || !CommDestroy || !GetErrorString) {
            err = "libnccl lacks a required symbol"; return false;
        }
        return true;
    }
};
NcclApi g_nccl;

template <typename T>
struct PinnedBuf {                 // growable page-locked staging buffer
    T *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)64);
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

template <typename T>
struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)64);
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct ProfSlot {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;   // recorded launches
    size_t used = 0;
    int64_t launches = 0;
    double bytes = 0.0, ms = 0.0;
};

// Staging workers: compiling ~10^3 random trees per generation is branch-misprediction-bound on one core
// (random trees = unpredictable control flow), so gsgp_generation spreads the trees over a few persistent
// threads; each writes its programs into its own buffer, the caller concatenates.
class StagePool {
public:
    struct Job {
        const int32_t *pool = nullptr, *off = nullptr, *len = nullptr;
        int32_t pool_len = 0, n_symbols = 0, t0 = 0, t1 = 0;
        std::vector<Ins> out;                 // programs of slots [t0, t1), concatenated
        std::vector<int32_t> rel_off, desc;   // per slot: offset inside `out`, prog_desc
        std::vector<uint8_t> need;
        std::vector<TreeFrame> frames;
        int32_t stack_need = 0, bad_slot = -1;
        const char *err = "";
        void run()
        {
            out.clear(); rel_off.assign((size_t)(t1 - t0), 0); desc.assign((size_t)(t1 - t0), 0);
            stack_need = 0; bad_slot = -1;
            size_t total = 1;
            for (int32_t t = t0; t < t1; ++t) if (len[t] > 0 && off[t] >= 0) total += (size_t)len[t] / 4 + 1;
            out.resize(total);
            size_t used = 0;
            for (int32_t t = t0; t < t1; ++t) {
                if (len[t] <= 0 || off[t] < 0) continue;
                if ((int64_t)off[t] + len[t] > pool_len) { bad_slot = t; err = "tree exceeds the tree pool"; return; }
                int n_ins = 0, nd = 0;
                if (!compile_prefix_tree(pool + off[t], len[t], n_symbols, out.data() + used, n_ins, nd, need, frames, err)) { bad_slot = t; return; }
                rel_off[(size_t)(t - t0)] = (int32_t)used; desc[(size_t)(t - t0)] = prog_desc(n_ins, nd);
                used += (size_t)n_ins;
                if (nd > stack_need) stack_need = nd;
            }
            out.resize(used);
        }
    };
    explicit StagePool(int n_threads) : jobs_((size_t)std::max(1, n_threads))
    {
        for (size_t i = 1; i < jobs_.size(); ++i) threads_.emplace_back([this, i] { worker(i); });
    }
    ~StagePool()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++epoch_; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    std::vector<Job> &jobs() { return jobs_; }
    void run_caller_only() { jobs_[0].run(); }       // small generations: waking the workers costs more than the work
    void run_all()                                   // job 0 on the caller, the others on the workers
    {
        { std::lock_guard<std::mutex> lk(m_); pending_ = (int)jobs_.size() - 1; ++epoch_; }
        cv_.notify_all();
        jobs_[0].run();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
private:
    void worker(size_t i)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            jobs_[i].run();
            { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::vector<Job> jobs_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    uint64_t epoch_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace

constexpr int kDrawSets = 8;                                  // draws run kDrawSets-1 generations ahead ...
constexpr int kDrawStreams = 6;                               // ... up to six of them at a time
constexpr int kRunHistBlock = 64;

struct gsgp_ctx {
    gsgp_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    Layout L{};
    int32_t tr_lo = 0, tr_hi = 0, te_lo = 0, te_hi = 0;     // global row ranges of this shard
    int32_t n_symbols = 0;
    int32_t n_tiles = 0;                                      // 2048-case tiles of gsop_kernel
    int32_t n_tiles_gen = 0;                                  // 256-case tiles of gen_kernel
    bool linsc = false;                                       // GSGP_FLAG_LINEAR_SCALING
    double tbar = 0.0;                                        // mean of the (global) train targets
    double *dsums2 = nullptr, *dsums_all2 = nullptr, *dab = nullptr;   // linear scaling: moment sums, a|b per individual
    bool fused = false;                                       // generation = one gen_kernel launch per chunk
    int32_t gen_levels = 2;
    int32_t chunk_inds = 0;                                   // individuals per interp+gsop launch pair
    int32_t tree_stride = 0, prog_stride = 0;
    int32_t generation = 0;
    int cur = 0;                                              // which store buffer holds generation g
    bool have_dataset = false, fitness_valid = false;

    double *dX = nullptr, *dy = nullptr, *dsym = nullptr;
    double *dsem[2] = {nullptr, nullptr};                     // [P][pitch] each; update_tables flips `cur`
    // single-buffer store (GSGP_FLAG_SINGLE_BUFFER_STORE, or automatically when two buffers do not fit): ONE allocation
    // [P][n_pad + shift]; dsem[0] = base, dsem[1] = base + shift.  A generation moves every row by `shift` cases
    // (alternately right and left), one block of `shift` cases per launch in the order that only ever overwrites
    // cases whose old values every individual has already consumed.  Memory = (1 + shift/n_pad) x the store, no copies.
    bool single_store = false;
    int32_t gen_warps = 16;                         // warps per gen_kernel block
    int32_t gen_waves = 8;                          // blocks per SM a launch aims for (GSGP_GEN_WAVES)
    int32_t shift_tiles = 0;                                  // case tiles (256 cases) per block
    int64_t pitch = 0;                                        // row pitch of the store in doubles
    double *dfit[2] = {nullptr, nullptr}, *dfit_test[2] = {nullptr, nullptr};
    double *dR = nullptr;
    double *dpartial = nullptr, *dsums = nullptr, *dsums_all = nullptr;
    unsigned char *dpack = nullptr; size_t pack_bytes = 0;        // one allocation: dplan | dprog_off | ds[0].prog_desc
    PinnedBuf<unsigned char> hpack;                               // its pinned image (gsgp_generation_compiled)
    gsgp_plan_entry *dplan = nullptr;                             // the plan of the generation being (or last) computed
    gsgp_plan_entry *dplan_buf[2] = {nullptr, nullptr};           // device mode: generation g's plan lives in buffer g & 1
    int32_t preselected = -1;                                     // generation whose selection rode in the previous finalize launch
    // random trees + programs of a generation.  Host replay uses set 0; device mode rotates through three sets:
    // generation g is consumed from set g%3 while plan_draw_kernel fills the sets of g+1 and g+2 on the second
    // stream (and the set of g stays readable for gsgp_get_tree until the draws of g+3).
    struct DrawSet {
        int32_t *tree_pool = nullptr;                         // device mode: reference arrays (gsgp_get_tree)
        DevBuf<Ins> prog_pool;                                // accumulator programs (program.hpp)
        int32_t *prog_desc = nullptr;                         // [2P] prog_desc(n, need)
        DrawRec *rec = nullptr;                               // device mode: what plan_draw_kernel drew
        int32_t *cand = nullptr;                              // device mode: tournament candidates
    } ds[kDrawSets];
    int dcur = 0;                                             // set of the current / last generation
    int32_t drawn[kDrawSets];                                 // generation whose draws sit in (or are on their way to) ds[i]
    cudaStream_t stream2 = nullptr;                           // plan_draw_kernel (= draw_stream[0])
    cudaStream_t draw_stream[kDrawStreams] = {};              // draws of consecutive generations overlap each other
    bool draw_concurrent = false;                             // the draw kernel keeps its scratch in shared memory
    cudaEvent_t ev_draw[kDrawSets] = {}, ev_free[kDrawSets] = {};   // per draw set: filled / no longer read
    uint16_t *dpost_pool = nullptr;                           // device mode: postfix scratch of plan_draw_kernel
    uint32_t *dcompile_scratch = nullptr;                     // device mode: program-builder scratch
    int32_t *dprog_off = nullptr;
    int32_t post_stride = 0;
    int32_t interp_tile = 256;                                // tunable (GSGP_INTERP_TILE)
    int sm_count = 148;
    int32_t *dindex_best = nullptr;
    BestRec *dhist = nullptr; int32_t hist_cap = 0;
    // GSGP_FLAG_KEEP_RUN_HISTORY: per generation, in blocks of kRunHistBlock generations
    bool keep_run = false;
    std::vector<double *> run_sem;                            // [block] -> kRunHistBlock rows of n_pad doubles
    std::vector<gsgp_plan_entry *> run_plan;                  // [block] -> kRunHistBlock plans of P entries

    PinnedBuf<gsgp_plan_entry> hplan;
    PinnedBuf<Ins> hprog;
    PinnedBuf<int32_t> hoff, hlen;
    PinnedBuf<double> hfit;
    PinnedBuf<int32_t> hint;
    // staging scratch reused across calls (no allocation per tree / per generation on the replay path)
    std::vector<uint8_t> st_need;
    std::vector<TreeFrame> st_tframes;
    std::unique_ptr<StagePool> stage_pool;                    // replay mode, created on first use
    std::vector<Ins> st_progs;
    std::vector<int32_t> st_off, st_len;
    std::vector<int32_t> last_pool;                           // replay mode: last generation's trees
    std::vector<gsgp_plan_entry> last_plan;

    ncclComm_t comm = nullptr;
    // peer exchange of the partial sums (kernels.cuh, reduce_partials_kernel): CUDA IPC mappings of every rank's
    // receive buffer; falls back to ncclAllGather when the mappings cannot be opened (or GSGP_EXCHANGE=nccl)
    bool p2p = false;
    unsigned char *xbase = nullptr;                           // local: [kXchgSlots][world][2P] doubles | flags | counter | err
    void *xpeer_base[kMaxPeers] = {};                         // opened mappings of the peers' xbase
    XchgDev xdev{};
    size_t xflags_off = 0;
    int32_t *xerr = nullptr;
    unsigned long long xseq = 0;
    ProfSlot prof[GSGP_K_COUNT];
    int profiling = 0;                                        // 0 off, 1 generation kernels only, 2 every kernel
    int64_t launch_count = 0;
};

namespace {

int fail(gsgp_ctx *c, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return 1;
}

#define CU(c, call)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail((c), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline int32_t round_up(int32_t x, int32_t m) { return (x + m - 1) / m * m; }

struct ProfScope {                 // CUDA-event bracket around one kernel launch on ctx->stream
    gsgp_ctx *c; int k; bool rec = false; size_t idx = 0; cudaStream_t st;
    ProfScope(gsgp_ctx *c_, int k_, double bytes, cudaStream_t st_ = nullptr) : c(c_), k(k_), st(st_ ? st_ : c_->stream)
    {
        c->launch_count++;
        ProfSlot &s = c->prof[k];
        s.launches++;
        // timing events cost ~2 us of stream time each: level 1 brackets only the kernels that matter for the roofline
        if (!c->profiling || (c->profiling == 1 && k != GSGP_K_GSOP && k != GSGP_K_INTERP)) return;
        if (s.used == s.ev.size()) {
            if (s.ev.size() >= 16384) return;
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            s.ev.push_back({a, b});
        }
        idx = s.used++; rec = true; s.bytes += bytes;
        cudaEventRecord(s.ev[idx].first, st);
    }
    ~ProfScope() { if (rec) cudaEventRecord(c->prof[k].ev[idx].second, st); }
};

// ---- launches ---------------------------------------------------------------------------------

// launch geometry of the interpreter: case tile (staged in shared memory when the variable columns of
// a tile fit), tree-slot groups so that the grid covers the SMs several times over
constexpr int32_t kMaxGroupSlots = 1024;
struct InterpGeom { int32_t tile; bool stage; size_t smem; int32_t groups, slots_per_group; };

InterpGeom interp_geometry(const gsgp_ctx *c, int32_t n_slots, int32_t levels)
{
    InterpGeom g{};
    const size_t fixed = (size_t)kInterpWarps * 2 * kProgSmem * sizeof(Ins) + 64;
    const size_t spill = (size_t)kInterpWarps * levels * kPass * sizeof(double);
    const size_t ncol = (size_t)c->cfg.nvar + c->cfg.num_constants;
    g.tile = 256; g.stage = false;
    for (int32_t t : {c->interp_tile, 256}) {
        if (ncol * t * sizeof(double) + spill + fixed <= (size_t)112 * 1024) { g.tile = t; g.stage = true; break; }
    }
    const int32_t n_tiles = (c->L.n_pad + g.tile - 1) / g.tile;
    const int32_t target = c->sm_count * 8;
    int32_t groups = std::max(1, (target + n_tiles - 1) / n_tiles);
    groups = std::min(groups, std::max(1, (n_slots + 15) / 16));
    groups = std::max(groups, (n_slots + kMaxGroupSlots - 1) / kMaxGroupSlots);   // slot table lives in shared memory
    g.slots_per_group = (n_slots + groups - 1) / groups;
    g.groups = (n_slots + g.slots_per_group - 1) / g.slots_per_group;
    g.smem = fixed + 8 * (size_t)g.slots_per_group + (g.stage ? ncol * g.tile * sizeof(double) + spill : 0);
    return g;
}

int launch_interp(gsgp_ctx *c, const Ins *prog_pool, const int32_t *prog_off, const int32_t *prog_desc,
                  int32_t first_slot, int32_t n_slots, int32_t stack_need, double *out, int64_t out_pitch,
                  double n_trees_hint)
{
    if (n_slots <= 0) return 0;
    if (stack_need > kMaxStackLevels) return fail(c, "tree needs %d spill levels > %d", stack_need, kMaxStackLevels);
    // spill levels kept in shared memory: what the launch needs, at most kMaxSmemLevels (deeper trees,
    // ~1 in 10^4 with Sethi-Ullman ordering, take the generic path)
    const int32_t levels = std::max(1, std::min(stack_need, kMaxSmemLevels));
    const InterpGeom g = interp_geometry(c, n_slots, levels);
    if (g.groups > 65535) return fail(c, "too many tree slots in one launch (%d)", n_slots);
    InterpArgs a{};
    a.prog_pool = prog_pool; a.prog_off = prog_off; a.prog_desc = prog_desc;
    a.first_slot = first_slot; a.n_slots = n_slots; a.slots_per_group = g.slots_per_group;
    a.X = c->dX; a.sym_val = c->dsym; a.nvar = c->cfg.nvar; a.nconst = c->cfg.num_constants;
    a.out = out; a.out_pitch = out_pitch; a.L = c->L;
    a.tile = g.tile; a.smem_levels = levels;
    dim3 grid((c->L.n_pad + g.tile - 1) / g.tile, g.groups);
    ProfScope ps(c, GSGP_K_INTERP, n_trees_hint * 8.0 * (double)(c->L.n_tr + c->L.n_te));
    if (g.stage) interp_kernel<true><<<grid, kInterpThreads, g.smem, c->stream>>>(a);
    else interp_kernel<false><<<grid, kInterpThreads, g.smem, c->stream>>>(a);
    CU(c, cudaGetLastError());
    return 0;
}

int launch_gsop(gsgp_ctx *c, int mode, int32_t k0, int32_t nk, int32_t r_k0, const double *sem_old,
                double *sem_new, double bytes)
{
    if (nk <= 0) return 0;
    GsArgs a{};
    a.plan = c->dplan; a.sem_old = sem_old; a.sem_new = sem_new; a.R = c->dR; a.y = c->dy;
    a.partial = c->dpartial; a.pitch = c->pitch; a.r_pitch = c->L.n_pad; a.k0 = k0; a.r_k0 = r_k0; a.n_tiles = c->n_tiles;
    a.mode = mode; a.L = c->L;
    if (nk > 65535) return fail(c, "gsop_kernel: %d individuals in one launch (the grid's y dimension holds 65535)", nk);
    dim3 grid(c->n_tiles, nk);
    ProfScope ps(c, GSGP_K_GSOP, bytes);
    switch (c->linsc ? GSGP_SUMO : c->cfg.error_measure) {   // linear scaling: this pass only sums the outputs
    case GSGP_SUMO: gsop_kernel<GSGP_SUMO><<<grid, kGsThreads, 0, c->stream>>>(a); break;
    case GSGP_MAE: gsop_kernel<GSGP_MAE><<<grid, kGsThreads, 0, c->stream>>>(a); break;
    case GSGP_MRE: gsop_kernel<GSGP_MRE><<<grid, kGsThreads, 0, c->stream>>>(a); break;
    default:       gsop_kernel<GSGP_MSE><<<grid, kGsThreads, 0, c->stream>>>(a); break;
    }
    CU(c, cudaGetLastError());
    return 0;
}

// fused pipeline: shared memory of one gen_kernel block for `ipg` individuals per group
size_t gen_smem_bytes(const gsgp_ctx *c, int32_t levels, int32_t ipg, int32_t warps)
{
    const size_t ncol = (size_t)c->cfg.nvar + c->cfg.num_constants;
    const size_t per_warp = ((size_t)levels * kPass + 2 * 2 * kGenProgSlots) * sizeof(double);
    return (ncol + 1) * kPass * sizeof(double) + per_warp * warps + (size_t)ipg * (sizeof(GenSlot) + 8) + sizeof(uint64_t) + 16 +
           4 * kCostBuckets;
}

// the instantiated block sizes of gen_kernel (warps per block, one block per SM)
#define GSGP_GEN_WARPS(X) X(28) X(24) X(20) X(16) X(12)

template <int MEASURE>
int launch_gen_kernel(gsgp_ctx *c, dim3 grid, size_t smem, const GenArgs &a)
{
#define GSGP_LAUNCH(W) if (c->gen_warps == W) { gen_kernel<MEASURE, W><<<grid, W * 32, smem, c->stream>>>(a); return 0; }
    GSGP_GEN_WARPS(GSGP_LAUNCH)
#undef GSGP_LAUNCH
    return fail(c, "gen_kernel: %d warps per block is not instantiated", c->gen_warps);
}

template <int MEASURE>
cudaError_t gen_kernels_allow_smem(int bytes)
{
    cudaError_t e = cudaSuccess;
#define GSGP_ATTR(W) if (e == cudaSuccess) e = cudaFuncSetAttribute(gen_kernel<MEASURE, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    GSGP_GEN_WARPS(GSGP_ATTR)
#undef GSGP_ATTR
    return e;
}

int launch_gen_tiles(gsgp_ctx *c, int32_t k0, int32_t nk, int32_t slot_k0, const double *sem_old, double *sem_new,
                     int32_t tile_begin, int32_t tile_count, double bytes)
{
    if (nk <= 0 || tile_count <= 0) return 0;
    const int32_t n_blocks = tile_count;
    const int32_t waves = c->gen_waves;
    const int32_t target = c->sm_count * waves;
    int32_t groups = std::max(1, (target + n_blocks - 1) / n_blocks);
    const int32_t groups_max = std::max(1, (nk + c->gen_warps - 1) / c->gen_warps);   // >= one individual per warp
    groups = std::min(groups, groups_max);
    // (The group count is not tuned to the launch's last "round" of blocks: 391 tiles x 3 groups fill 7.93 rounds of 148
    // SMs and x 4 groups 10.57, yet 4 groups are faster, 0.593 against 0.628 ms -- blocks end at different times, and
    // fewer, longer blocks lengthen the launch's tail.)
    int32_t ipg = (nk + groups - 1) / groups;
    while (ipg > 32 && gen_smem_bytes(c, c->gen_levels, ipg, c->gen_warps) > (size_t)227 * 1024) ipg = (ipg + 1) / 2;
    groups = (nk + ipg - 1) / ipg;
    if (groups > 65535) return fail(c, "population chunk too large for one launch");
    if (c->gen_levels < 2) return fail(c, "gen_kernel needs two spill levels (sigma(R2) of a mutation waits in the last one)");
    GenArgs a{};
    a.plan = c->dplan; a.prog_pool = c->ds[c->dcur].prog_pool.p; a.prog_off = c->dprog_off; a.prog_desc = c->ds[c->dcur].prog_desc;
    a.slot_k0 = slot_k0; a.X = c->dX; a.y = c->dy; a.sym_val = c->dsym; a.nvar = c->cfg.nvar; a.nconst = c->cfg.num_constants;
    a.sem_old = sem_old; a.sem_new = sem_new; a.pitch = c->pitch; a.tile_begin = tile_begin;
    a.partial = c->dpartial;
    a.k0 = k0; a.nk = nk; a.inds_per_group = ipg; a.n_tiles = c->n_tiles_gen; a.smem_levels = c->gen_levels; a.L = c->L;
    const size_t smem = gen_smem_bytes(c, c->gen_levels, ipg, c->gen_warps);
    dim3 grid(n_blocks, groups);
    ProfScope ps(c, GSGP_K_GSOP, bytes);
    int rc;
    switch (c->linsc ? GSGP_SUMO : c->cfg.error_measure) {
    case GSGP_SUMO: rc = launch_gen_kernel<GSGP_SUMO>(c, grid, smem, a); break;
    case GSGP_MAE: rc = launch_gen_kernel<GSGP_MAE>(c, grid, smem, a); break;
    case GSGP_MRE: rc = launch_gen_kernel<GSGP_MRE>(c, grid, smem, a); break;
    default:       rc = launch_gen_kernel<GSGP_MSE>(c, grid, smem, a); break;
    }
    if (rc) return rc;
    CU(c, cudaGetLastError());
    return 0;
}

// one generation's fused launch(es) over individuals [k0, k0+nk): straight into the other store buffer, or -- single-
// buffer store -- block of cases by block, each launch writing where the previous one has finished reading
int launch_gen(gsgp_ctx *c, int32_t k0, int32_t nk, int32_t slot_k0, const double *sem_old, double *sem_new, double bytes)
{
    if (!c->single_store)
        return launch_gen_tiles(c, k0, nk, slot_k0, sem_old, sem_new, 0, c->n_tiles_gen, bytes);
    if (k0 != 0 || nk != c->cfg.population_size) return fail(c, "single-buffer store: a generation is one launch sequence");
    const int32_t nb = (c->n_tiles_gen + c->shift_tiles - 1) / c->shift_tiles;
    const bool right = sem_new > sem_old;                      // rows move right: last block first (it lands in the free tail)
    for (int32_t i = 0; i < nb; ++i) {
        const int32_t b = right ? nb - 1 - i : i;
        const int32_t t0 = b * c->shift_tiles, nt = std::min(c->shift_tiles, c->n_tiles_gen - t0);
        if (launch_gen_tiles(c, k0, nk, slot_k0, sem_old, sem_new, t0, nt, bytes * nt / c->n_tiles_gen)) return 1;
    }
    return 0;
}

// local partials -> (all-gather across ranks) -> fitness -> best-of-generation
// tile partials -> per-individual local sums -> every rank's sums (rank order is the summation order)
int reduce_and_gather(gsgp_ctx *c, int mode, int32_t n_tiles, double *local, double *gathered, const double **all)
{
    const int P = c->cfg.population_size;
    const int W = c->cfg.world_size;
    ReduceArgs r{c->dplan, c->dpartial, local, P, n_tiles, mode, XchgDev{}};
    if (c->p2p) {                                                      // fused: sums go straight into every peer's slot
        r.x = c->xdev;
        r.x.seq = ++c->xseq;
        r.x.slot = (int32_t)(c->xseq % kXchgSlots);
    }
    {
        ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
        reduce_partials_kernel<<<(P * 32 + 255) / 256, 256, 0, c->stream>>>(r);
        CU(c, cudaGetLastError());
    }
    *all = local;
    if (c->p2p) {
        static const unsigned long long timeout_ns =
            (unsigned long long)(getenv("GSGP_EXCHANGE_TIMEOUT_MS") ? atof(getenv("GSGP_EXCHANGE_TIMEOUT_MS")) : 20000.0) * 1000000ull;
        const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(c->xbase + c->xflags_off) + (size_t)r.x.slot * W;
        ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
        xchg_wait_kernel<<<1, 32, 0, c->stream>>>(flags, W, r.x.seq, timeout_ns, c->xerr);
        CU(c, cudaGetLastError());
        *all = reinterpret_cast<const double *>(c->xbase) + (size_t)r.x.slot * W * 2 * P;
    } else if (W > 1) {
        ncclResult_t rc = g_nccl.AllGather(local, gathered, (size_t)2 * P, ncclDouble, c->comm, c->stream);
        if (rc != ncclSuccess) return fail(c, "ncclAllGather failed: %s", g_nccl.GetErrorString(rc));
        *all = gathered;
    }
    return 0;
}

// a byte-wise all-gather through NCCL on the context's stream (setup-time exchanges and barriers)
int nccl_gather_bytes(gsgp_ctx *c, const void *mine, void *all_host, size_t bytes)
{
    const int W = c->cfg.world_size;
    unsigned char *d = nullptr;
    CU(c, cudaMalloc((void **)&d, bytes * (W + 1)));
    CU(c, cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, c->stream));
    ncclResult_t rc = g_nccl.AllGather(d, d + bytes, bytes, ncclChar, c->comm, c->stream);
    if (rc != ncclSuccess) { cudaFree(d); return fail(c, "ncclAllGather failed: %s", g_nccl.GetErrorString(rc)); }
    CU(c, cudaMemcpyAsync(all_host, d + bytes, bytes * W, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

// maps every rank's receive buffer into this process (CUDA IPC over NVLink peer access).  All ranks agree on the
// outcome: one rank that cannot open a mapping sends everybody back to the NCCL all-gather.
int setup_peer_exchange(gsgp_ctx *c)
{
    const int W = c->cfg.world_size, P = c->cfg.population_size;
    const char *ex = getenv("GSGP_EXCHANGE");
    bool want = W <= kMaxPeers && !(ex && strcmp(ex, "nccl") == 0);
    const size_t data_bytes = sizeof(double) * kXchgSlots * (size_t)W * 2 * P;
    c->xflags_off = (data_bytes + 255) / 256 * 256;
    const size_t flag_bytes = sizeof(unsigned long long) * kXchgSlots * (size_t)W;
    const size_t ctr_off = c->xflags_off + (flag_bytes + 255) / 256 * 256;
    const size_t total = ctr_off + 256;
    cudaIpcMemHandle_t mine{};
    unsigned char ok = 0;
    if (want) {
        CU(c, cudaMalloc((void **)&c->xbase, total));
        CU(c, cudaMemsetAsync(c->xbase, 0, total, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        ok = cudaIpcGetMemHandle(&mine, c->xbase) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    std::vector<cudaIpcMemHandle_t> handles((size_t)W);
    if (nccl_gather_bytes(c, &mine, handles.data(), sizeof(mine))) return 1;
    std::vector<unsigned char> oks((size_t)W);
    if (nccl_gather_bytes(c, &ok, oks.data(), 1)) return 1;
    bool all_ok = want;
    for (unsigned char v : oks) all_ok = all_ok && v;
    {   // one rank per GPU is a hard requirement of the peer-store exchange: xchg_wait_kernel spins on flags that the
        // other ranks' reduce kernels raise, and two ranks sharing a GPU could keep each other's kernels off the SMs
        // (B200_PROFILING.md: cross-launch spin-waits on one GPU end in a context-switch timeout).  Ranks on the same
        // device (same UUID) therefore take the ncclAllGather path.
        cudaDeviceProp prop{};
        CU(c, cudaGetDeviceProperties(&prop, c->cfg.device));
        std::vector<cudaUUID_t> ids((size_t)W);
        if (nccl_gather_bytes(c, &prop.uuid, ids.data(), sizeof(cudaUUID_t))) return 1;
        for (int r = 0; r < W && all_ok; ++r)
            for (int q = r + 1; q < W; ++q)
                if (memcmp(&ids[(size_t)r], &ids[(size_t)q], sizeof(cudaUUID_t)) == 0) { all_ok = false; break; }
    }
    unsigned char opened = all_ok;
    if (all_ok) {
        for (int r = 0; r < W && opened; ++r) {
            if (r == c->cfg.rank) { c->xpeer_base[r] = nullptr; continue; }
            if (cudaIpcOpenMemHandle(&c->xpeer_base[r], handles[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError(); c->xpeer_base[r] = nullptr; opened = 0;
            }
        }
    }
    if (nccl_gather_bytes(c, &opened, oks.data(), 1)) return 1;     // also the barrier: every buffer is zeroed and mapped
    bool use = all_ok;
    for (unsigned char v : oks) use = use && v;
    if (!use) {
        for (int r = 0; r < W; ++r) if (c->xpeer_base[r]) { cudaIpcCloseMemHandle(c->xpeer_base[r]); c->xpeer_base[r] = nullptr; }
        cudaFree(c->xbase); c->xbase = nullptr;
        c->p2p = false;
        return 0;
    }
    for (int r = 0; r < W; ++r) {
        unsigned char *b = r == c->cfg.rank ? c->xbase : static_cast<unsigned char *>(c->xpeer_base[r]);
        c->xdev.peer_data[r] = reinterpret_cast<double *>(b);
        c->xdev.peer_flags[r] = reinterpret_cast<unsigned long long *>(b + c->xflags_off);
    }
    c->xdev.counter = reinterpret_cast<unsigned int *>(c->xbase + ctr_off);
    c->xerr = reinterpret_cast<int32_t *>(c->xbase + ctr_off + 64);
    c->xdev.world = W; c->xdev.rank = c->cfg.rank;
    c->p2p = true;
    return 0;
}

template <int PASS>
int launch_ls_pass(gsgp_ctx *c, int mode, const double *sem, const double *osum_all)
{
    const int P = c->cfg.population_size;
    LsArgs a{};
    a.plan = c->dplan; a.sem = sem; a.y = c->dy; a.partial = c->dpartial; a.osum_all = osum_all; a.ab = c->dab;
    a.pitch = c->pitch; a.P = P; a.world = c->cfg.world_size; a.n_tiles = c->n_tiles; a.mode = mode;
    a.nrow = c->cfg.nrow; a.tbar = c->tbar; a.L = c->L;
    const double n_cases = (double)(c->L.n_tr + c->L.n_te);
    if (P > 65535) return fail(c, "linear scaling supports populations up to 65535");
    dim3 grid(c->n_tiles, P);
    ProfScope ps(c, GSGP_K_GSOP, 8.0 * P * n_cases);
    if (PASS == LS_MOMENTS) ls_pass_kernel<LS_MOMENTS, GSGP_MSE><<<grid, kGsThreads, 0, c->stream>>>(a);
    else switch (c->cfg.error_measure) {
        case GSGP_MAE: ls_pass_kernel<LS_ERROR, GSGP_MAE><<<grid, kGsThreads, 0, c->stream>>>(a); break;
        case GSGP_MRE: ls_pass_kernel<LS_ERROR, GSGP_MRE><<<grid, kGsThreads, 0, c->stream>>>(a); break;
        default:       ls_pass_kernel<LS_ERROR, GSGP_MSE><<<grid, kGsThreads, 0, c->stream>>>(a); break;
    }
    CU(c, cudaGetLastError());
    return 0;
}

// local partials -> (all-gather across ranks) -> fitness -> best-of-generation.  n_tiles: tiling of the partials
// the generation kernels left (gsop: 2048-case tiles, gen_kernel: 256-case tiles)
int ensure_hist(gsgp_ctx *c, int32_t gen);

PlanArgs plan_args(gsgp_ctx *c, int32_t generation);

// best_gen >= 0: best_individual of the new fitness rides in the same launch (history slot best_gen)
int launch_finalize(gsgp_ctx *c, int mode, int old_buf, int new_buf, int32_t n_tiles, int32_t best_gen = -1)
{
    const int P = c->cfg.population_size;
    const double *sums_all = nullptr;
    if (reduce_and_gather(c, mode, n_tiles, c->dsums, c->dsums_all, &sums_all)) return 1;
    if (c->linsc) {
        // -linsc (main_cpu.go:991-1104,1142-1177): the sums above are the train outputs' sums; centred moments, a and
        // b, then the error of a + b*o over the new rows
        const double *osum_all = sums_all, *mom_all = nullptr;
        if (launch_ls_pass<LS_MOMENTS>(c, mode, c->dsem[new_buf], osum_all)) return 1;
        if (reduce_and_gather(c, mode, c->n_tiles, c->dsums2, c->dsums_all2, &mom_all)) return 1;
        {
            ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
            ls_ab_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(osum_all, mom_all, c->dab, P, c->cfg.world_size, c->cfg.nrow, c->tbar);
            CU(c, cudaGetLastError());
        }
        if (launch_ls_pass<LS_ERROR>(c, mode, c->dsem[new_buf], osum_all)) return 1;
        if (reduce_and_gather(c, mode, c->n_tiles, c->dsums, c->dsums_all, &sums_all)) return 1;
    }
    {
        FitnessArgs f{c->dplan, sums_all, c->dfit[old_buf], c->dfit_test[old_buf], c->dfit[new_buf],
                      c->dfit_test[new_buf], P, c->cfg.world_size, mode, c->cfg.error_measure,
                      c->cfg.nrow, c->cfg.nrow_test};
        if (best_gen >= 0 && ensure_hist(c, best_gen)) return 1;
        // device mode: the next generation's selection rides in the same launch (its draws run generations ahead)
        static const bool no_fold = getenv("GSGP_NO_FOLDED_FINALIZE") != nullptr;
        const int32_t next_gen = best_gen + 1;
        const bool fold = !no_fold && best_gen >= 1 && mode == GS_MODE_GENERATION && c->cfg.rng_mode == GSGP_RNG_DEVICE_PHILOX &&
                          c->drawn[next_gen % kDrawSets] == next_gen;
        PlanArgs next{};
        if (fold) {
            next = plan_args(c, next_gen);
            next.fit = c->dfit[new_buf];
            CU(c, cudaStreamWaitEvent(c->stream, c->ev_draw[next_gen % kDrawSets], 0));
        }
        ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
        if (fold) {
            finalize_select_kernel<<<1, 1024, 0, c->stream>>>(ReduceArgs{}, f, c->cfg.minimization_problem, c->dindex_best, c->dhist + best_gen, next, 0, 1);
            c->preselected = next_gen;
        } else if (best_gen >= 0)
            fitness_best_kernel<<<1, 1024, 0, c->stream>>>(f, c->cfg.minimization_problem, c->dindex_best, c->dhist + best_gen);
        else fitness_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(f);
        CU(c, cudaGetLastError());
    }
    return 0;
}

int ensure_hist(gsgp_ctx *c, int32_t gen)
{
    if (gen < c->hist_cap) return 0;
    int32_t cap = std::max(gen + 1, std::max(1024, c->hist_cap * 2));
    BestRec *n = nullptr;
    CU(c, cudaMalloc((void **)&n, sizeof(BestRec) * (size_t)cap));
    CU(c, cudaMemsetAsync(n, 0, sizeof(BestRec) * (size_t)cap, c->stream));
    if (c->dhist) {
        CU(c, cudaMemcpyAsync(n, c->dhist, sizeof(BestRec) * (size_t)c->hist_cap, cudaMemcpyDeviceToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->dhist);
    }
    c->dhist = n; c->hist_cap = cap;
    return 0;
}

int ensure_run_history(gsgp_ctx *c, int32_t gen)
{
    while ((int32_t)c->run_sem.size() * kRunHistBlock <= gen) {
        double *s = nullptr; gsgp_plan_entry *p = nullptr;
        CU(c, cudaMalloc((void **)&s, sizeof(double) * (size_t)kRunHistBlock * c->L.n_pad));
        CU(c, cudaMalloc((void **)&p, sizeof(gsgp_plan_entry) * (size_t)kRunHistBlock * c->cfg.population_size));
        c->run_sem.push_back(s); c->run_plan.push_back(p);
    }
    return 0;
}

int launch_best(gsgp_ctx *c, int buf, int32_t gen, bool best_done = false)
{
    if (ensure_hist(c, gen)) return 1;
    if (!best_done) {
        ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
        best_kernel<<<1, 1024, 0, c->stream>>>(c->dfit[buf], c->dfit_test[buf], c->cfg.population_size,
                                               c->cfg.minimization_problem, c->dindex_best, c->dhist + gen,
                                               reinterpret_cast<int32_t *>(c->dfit[buf] + 2 * (size_t)c->cfg.population_size));
        CU(c, cudaGetLastError());
    }
    if (c->keep_run) {                                        // best-of-generation semantic + this generation's decisions
        if (ensure_run_history(c, gen)) return 1;
        double *dst = c->run_sem[(size_t)gen / kRunHistBlock] + (size_t)(gen % kRunHistBlock) * c->L.n_pad;
        ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
        copy_best_row_kernel<<<std::min(1024, (c->L.n_pad / 2 + 255) / 256), 256, 0, c->stream>>>(c->dsem[buf], c->pitch, c->L.n_pad,
                                                                                                  c->dindex_best, dst);
        CU(c, cudaGetLastError());
        if (gen > 0)
            CU(c, cudaMemcpyAsync(c->run_plan[(size_t)gen / kRunHistBlock] + (size_t)(gen % kRunHistBlock) * c->cfg.population_size,
                                  c->dplan, sizeof(gsgp_plan_entry) * (size_t)c->cfg.population_size, cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

// after the generation kernels: fitness from the partials, update_tables, best_individual
PlanArgs plan_args(gsgp_ctx *c, int32_t generation);

// Small shapes on one GPU: the reduction, fitness + best individual AND (device mode) the next generation's selection in
// one single-block launch (finalize_select_kernel) -- the generation's chain of dependent launches is two kernels long
// instead of four.  Needs the next generation's draws (they run several generations ahead on their own streams).
bool small_shape(const gsgp_ctx *c, int32_t n_tiles)
{
    static const bool off = getenv("GSGP_NO_FOLDED_FINALIZE") != nullptr;
    return !off && c->cfg.world_size == 1 && !c->linsc && !c->keep_run &&
           c->cfg.population_size <= 4096 && (int64_t)c->cfg.population_size * n_tiles <= (1 << 16);
}

int launch_finalize_select(gsgp_ctx *c, int old_buf, int new_buf, int32_t n_tiles, int32_t gen)
{
    const int P = c->cfg.population_size;
    if (ensure_hist(c, gen)) return 1;
    ReduceArgs r{c->dplan, c->dpartial, c->dsums, P, n_tiles, GS_MODE_GENERATION, XchgDev{}};
    FitnessArgs f{c->dplan, c->dsums, c->dfit[old_buf], c->dfit_test[old_buf], c->dfit[new_buf], c->dfit_test[new_buf],
                  P, 1, GS_MODE_GENERATION, c->cfg.error_measure, c->cfg.nrow, c->cfg.nrow_test};
    const int32_t next_gen = gen + 1;
    // host replay has no draws on the device (the plan comes with the next call): reduction + fitness + best only
    const bool drawn = c->cfg.rng_mode == GSGP_RNG_DEVICE_PHILOX && c->drawn[next_gen % kDrawSets] == next_gen;
    PlanArgs next{};
    if (drawn) { next = plan_args(c, next_gen); next.fit = c->dfit[new_buf]; }   // the fitness this launch is about to write
    if (drawn) CU(c, cudaStreamWaitEvent(c->stream, c->ev_draw[next_gen % kDrawSets], 0));
    ProfScope ps(c, GSGP_K_FINALIZE, 0.0);
    finalize_select_kernel<<<1, 1024, 0, c->stream>>>(r, f, c->cfg.minimization_problem, c->dindex_best, c->dhist + gen, next, 1, drawn ? 1 : 0);
    CU(c, cudaGetLastError());
    c->preselected = drawn ? next_gen : -1;
    return 0;
}

int finish_generation(gsgp_ctx *c)
{
    const int old_buf = c->cur, new_buf = c->cur ^ 1;
    const int32_t n_tiles = c->fused ? c->n_tiles_gen : c->n_tiles;
    if (small_shape(c, n_tiles)) {
        if (launch_finalize_select(c, old_buf, new_buf, n_tiles, c->generation + 1)) return 1;
    } else if (launch_finalize(c, GS_MODE_GENERATION, old_buf, new_buf, n_tiles, c->generation + 1)) return 1;
    c->cur = new_buf;                                                  // update_tables main_cpu.go:1191-1197
    c->generation++;
    if (launch_best(c, c->cur, c->generation, true)) return 1;         // :1492 (best_individual rode in the fitness launch)
    c->fitness_valid = true;
    return 0;
}

// algorithmic bytes of a fused launch over individuals [k0, k0+nk) of a host plan: XO 24 B, MUT 16 B, copy 16 B
double gen_bytes(const gsgp_plan_entry *plan, int32_t k0, int32_t nk, double n_cases)
{
    double b = 0.0;
    for (int32_t k = k0; k < k0 + nk; ++k) {
        const bool computed = !plan[k].elite && plan[k].op != GSGP_OP_REPR;
        b += (computed && plan[k].op == GSGP_OP_XO) ? 24.0 : 16.0;
    }
    return b * n_cases;
}

// the interp -> gsop pipeline of one generation, chunked so that R fits its workspace
int run_generation_kernels(gsgp_ctx *c, const gsgp_plan_entry *host_plan /* may be NULL */, int32_t max_sp)
{
    const int P = c->cfg.population_size;
    const int old_buf = c->cur, new_buf = c->cur ^ 1;
    const double n_cases = (double)(c->L.n_tr + c->L.n_te);
    for (int32_t k0 = 0; k0 < P; k0 += c->chunk_inds) {
        const int32_t nk = std::min(c->chunk_inds, P - k0);
        double n_trees = 1.2 * nk, frac_xo = 0.6, frac_mut = 0.3;     // expected mix when the plan lives on the device
        if (host_plan) {
            n_trees = 0; double nx = 0, nm = 0;
            for (int32_t k = k0; k < k0 + nk; ++k) {
                n_trees += (host_plan[k].tree0_len > 0) + (host_plan[k].tree1_len > 0);
                const bool computed = !host_plan[k].elite && host_plan[k].op != GSGP_OP_REPR;
                nx += computed && host_plan[k].op == GSGP_OP_XO; nm += computed && host_plan[k].op == GSGP_OP_MUT;
            }
            frac_xo = nx / nk; frac_mut = nm / nk;
        }
        if (c->fused) {
            // algorithmic bytes of the fused launch: XO 24 B (p1, p2, child), MUT 16 B (parent, child), copy 16 B
            const double bytes = (24.0 * frac_xo + 16.0 * frac_mut + 16.0 * (1.0 - frac_xo - frac_mut)) * nk * n_cases;
            if (launch_gen(c, k0, nk, 0, c->dsem[old_buf], c->dsem[new_buf], bytes)) return 1;
        } else {
            const double bytes = (32.0 * (frac_xo + frac_mut) + 16.0 * (1.0 - frac_xo - frac_mut)) * nk * n_cases;
            if (launch_interp(c, c->ds[c->dcur].prog_pool.p, c->dprog_off, c->ds[c->dcur].prog_desc, 2 * k0, 2 * nk, max_sp, c->dR,
                              c->L.n_pad, n_trees)) return 1;
            if (launch_gsop(c, GS_MODE_GENERATION, k0, nk, k0, c->dsem[old_buf], c->dsem[new_buf], bytes)) return 1;
        }
    }
    // device mode: the draw set of this generation is no longer read once the kernels above are done
    if (!host_plan && c->stream2) CU(c, cudaEventRecord(c->ev_free[c->dcur], c->stream));
    return finish_generation(c);
}

PlanArgs plan_args(gsgp_ctx *c, int32_t generation)
{
    gsgp_ctx::DrawSet &d = c->ds[generation % kDrawSets];
    PlanArgs a{};
    a.plan = c->dplan_buf[generation & 1]; a.rec = d.rec; a.cand = d.cand; a.tree_pool = d.tree_pool; a.post_pool = c->dpost_pool;
    a.compile_scratch = c->dcompile_scratch; a.prog_pool = d.prog_pool.p; a.prog_desc = d.prog_desc;
    a.post_stride = c->post_stride;
    a.fit = c->dfit[c->cur]; a.index_best = c->dindex_best;
    a.P = c->cfg.population_size; a.nvar = c->cfg.nvar; a.nconst = c->cfg.num_constants;
    a.max_depth = c->cfg.max_depth_creation; a.zero_depth = c->cfg.zero_depth;
    a.tournament_size = c->cfg.tournament_size; a.minimize = c->cfg.minimization_problem;
    a.tree_stride = c->tree_stride; a.prog_stride = c->prog_stride;
    a.generation = generation; a.p_xo = c->cfg.p_crossover; a.p_mut = c->cfg.p_mutation;
    a.seed = c->cfg.rng_seed;
    return a;
}

// everything of generation `generation` that consumes random numbers, on the second stream; it may only
// overwrite its draw set once the generation that last read it (generation - 2) is done: ev_free[set]
int launch_draw(gsgp_ctx *c, int32_t generation)
{
    const PlanArgs a = plan_args(c, generation);
    // the kernel is latency-bound (one thread builds an individual's trees), so at small populations it, not the
    // generation kernels, sets the pace: draws of consecutive generations go to different streams and overlap
    cudaStream_t st = c->draw_concurrent ? c->draw_stream[generation % kDrawStreams] : c->stream2;
    CU(c, cudaStreamWaitEvent(st, c->ev_free[generation % kDrawSets], 0));
    {
        ProfScope ps(c, GSGP_K_PLAN, 0.0, st);
        // working set of a thread: postfix + builder scratch, post_stride words each, in shared memory when a block of
        // >= 32 threads fits (depth <= 8); global scratch pools otherwise
        const size_t per_thread = 2 * (size_t)c->post_stride * sizeof(uint32_t);
        int nt = 0;
        for (int cand : {64, 32}) if (per_thread * cand <= (size_t)200 * 1024) { nt = cand; break; }
        if (nt) plan_draw_kernel<true><<<(a.P + nt - 1) / nt, nt, per_thread * nt, st>>>(a);
        else plan_draw_kernel<false><<<(a.P + 127) / 128, 128, 0, st>>>(a);
        CU(c, cudaGetLastError());
    }
    CU(c, cudaEventRecord(c->ev_draw[generation % kDrawSets], st));
    c->drawn[generation % kDrawSets] = generation;
    return 0;
}

// tournament winners + elite rule for `generation` on the main stream (needs the previous generation's fit[])
int launch_select(gsgp_ctx *c, int32_t generation)
{
    const PlanArgs a = plan_args(c, generation);
    c->dcur = generation % kDrawSets;
    c->dplan = a.plan;
    if (c->preselected == generation) { c->preselected = -1; return 0; }    // chosen by the previous generation's finalize launch
    c->preselected = -1;
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_draw[generation % kDrawSets], 0));
    ProfScope ps(c, GSGP_K_PLAN, 0.0);
    plan_select_kernel<<<(a.P + 255) / 256, 256, 0, c->stream>>>(a);
    CU(c, cudaGetLastError());
    return 0;
}

// host trees (reference array format) -> accumulator programs; returns the spill levels they need
int stage_programs(gsgp_ctx *c, int32_t n, const int32_t *pool, int32_t pool_len, const int32_t *off,
                   const int32_t *len, std::vector<Ins> &progs, int32_t *hoff, int32_t *hlen, int32_t &stack_need)
{
    stack_need = 0;
    const char *err = "";
    size_t total = 1;
    for (int32_t t = 0; t < n; ++t) if (len[t] > 0 && off[t] >= 0) total += (size_t)len[t] / 4 + 1;   // function nodes + 1
    progs.resize(total);
    size_t used = 0;
    for (int32_t t = 0; t < n; ++t) {
        if (len[t] <= 0 || off[t] < 0) { hoff[t] = 0; hlen[t] = 0; continue; }
        if ((int64_t)off[t] + len[t] > pool_len) return fail(c, "tree %d exceeds the tree pool", t);
        int n_ins = 0, need = 0;
        if (!compile_prefix_tree(pool + off[t], len[t], c->n_symbols, progs.data() + used, n_ins, need, c->st_need, c->st_tframes, err))
            return fail(c, "tree %d: %s", t, err);
        hoff[t] = (int32_t)used; hlen[t] = prog_desc(n_ins, need);
        used += (size_t)n_ins;
        stack_need = std::max(stack_need, need);
    }
    progs.resize(used);
    return 0;
}

int check_ready(gsgp_ctx *c)
{
    if (!c) return fail(nullptr, "null context");
    if (!c->have_dataset) return fail(c, "dataset not set (gsgp_set_dataset)");
    CU(c, cudaSetDevice(c->cfg.device));
    return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int32_t gsgp_abi_version(void) { return GSGP_ABI_VERSION; }

void gsgp_config_defaults(gsgp_config *cfg)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->population_size = 200; cfg->max_depth_creation = 6; cfg->tournament_size = 4;
    cfg->minimization_problem = 1; cfg->error_measure = GSGP_MSE;
    cfg->p_crossover = 0.6; cfg->p_mutation = 0.3; cfg->rng_seed = 1;
    cfg->rng_mode = GSGP_RNG_HOST_REPLAY; cfg->world_size = 1;
}

const char *gsgp_last_error(const gsgp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gsgp_nccl_unique_id(uint8_t out[128])
{
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, "%s", err.c_str());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclResult_t rc = g_nccl.GetUniqueId(&id);
    if (rc != ncclSuccess) return fail(nullptr, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
    memcpy(out, &id, 128);
    return 0;
}

void gsgp_shard_range(int32_t n, int32_t rank, int32_t world, int32_t *lo, int32_t *hi)
{
    *lo = (int32_t)((int64_t)n * rank / world);
    *hi = (int32_t)((int64_t)n * (rank + 1) / world);
}

int gsgp_create(const gsgp_config *cfg, gsgp_ctx **out)
{
    if (!cfg || !out) return fail(nullptr, "null argument");
    *out = nullptr;
    if (cfg->population_size < 1 || cfg->population_size > 65535 * 16)
        return fail(nullptr, "population_size out of range");
    if (cfg->nvar < 1 || cfg->nrow < 0 || cfg->nrow_test < 0 || cfg->nrow + (int64_t)cfg->nrow_test < 1)
        return fail(nullptr, "bad dataset shape");
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return fail(nullptr, "bad rank/world_size");
    if (cfg->error_measure < GSGP_MSE || cfg->error_measure > GSGP_MRE) return fail(nullptr, "Unknown error measure");
    if (cfg->tournament_size < 1) return fail(nullptr, "tournament_size must be >= 1");
    if (cfg->p_crossover < 0 || cfg->p_mutation < 0 || cfg->p_crossover + cfg->p_mutation > 1)   // main_cpu.go:374-376
        return fail(nullptr, "Crossover rate and mutation rate must be greater or equal to 0 and their sum must be smaller or equal to 1.");
    if (4 + (int64_t)cfg->nvar + cfg->num_constants > kMaxSymbols) return fail(nullptr, "too many symbols");
    if (cfg->rng_mode == GSGP_RNG_DEVICE_PHILOX && (cfg->max_depth_creation < 1 || cfg->max_depth_creation > kMaxGenDepth))
        return fail(nullptr, "device RNG mode supports max_depth_creation in [1,%d]", kMaxGenDepth);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, "no CUDA device: the gsgp_b200 path has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "bad device ordinal %d", cfg->device);

    gsgp_ctx *c = new gsgp_ctx();
    c->cfg = *cfg;
    std::fill(c->drawn, c->drawn + kDrawSets, -1);
    auto bail = [&](void) { g_create_error = c->err; gsgp_destroy(c); return 1; };
#define CUC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fail(c, "%s failed: %s", #call, cudaGetErrorString(e__)); return bail(); } } while (0)
    CUC(cudaSetDevice(cfg->device));
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));

    gsgp_shard_range(cfg->nrow, cfg->rank, cfg->world_size, &c->tr_lo, &c->tr_hi);
    gsgp_shard_range(cfg->nrow_test, cfg->rank, cfg->world_size, &c->te_lo, &c->te_hi);
    c->L.n_tr = c->tr_hi - c->tr_lo; c->L.n_te = c->te_hi - c->te_lo;
    c->L.tr_pad = round_up(c->L.n_tr, 16);
    c->L.n_pad = c->L.tr_pad + round_up(c->L.n_te, 16);
    if (c->L.n_pad == 0) c->L.n_pad = 16;
    c->n_tiles = (c->L.n_pad + kGsTile - 1) / kGsTile;
    c->n_tiles_gen = (c->L.n_pad + kPass - 1) / kPass;
    c->n_symbols = 4 + cfg->nvar + cfg->num_constants;

    const int P = cfg->population_size;
    const size_t row = (size_t)c->L.n_pad;
    {
        cudaDeviceProp prop{};
        CUC(cudaGetDeviceProperties(&prop, cfg->device));
        c->sm_count = prop.multiProcessorCount;
        if (const char *e = getenv("GSGP_INTERP_TILE")) { int t = atoi(e); if (t >= kPass && t <= 4096 && t % kPass == 0) c->interp_tile = t; }
        const int kMaxSmem = 227 * 1024;
        CUC(cudaFuncSetAttribute(interp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CUC(cudaFuncSetAttribute(interp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        CUC(gen_kernels_allow_smem<GSGP_MSE>(kMaxSmem));
        CUC(gen_kernels_allow_smem<GSGP_MAE>(kMaxSmem));
        CUC(gen_kernels_allow_smem<GSGP_MRE>(kMaxSmem));
        CUC(gen_kernels_allow_smem<GSGP_SUMO>(kMaxSmem));
        CUC(cudaFuncSetAttribute(plan_draw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        c->linsc = (cfg->flags & GSGP_FLAG_LINEAR_SCALING) != 0;
        c->keep_run = (cfg->flags & GSGP_FLAG_KEEP_RUN_HISTORY) != 0;
        // pipeline: fused (one gen_kernel launch per generation, R stays on chip) unless the dataset's columns
        // do not fit shared memory, the caller wants R kept (proto dump), or GSGP_PIPELINE=unfused
        const char *pl = getenv("GSGP_PIPELINE");
        c->gen_levels = 2;
        // warps per gen_kernel block (one block per SM): the kernel is bound by its warps' issue rate, so as many as the
        // register file carries without spills (28 x 72 registers); fewer when the dataset's columns leave less shared
        // memory.  GSGP_GEN_WARPS picks one of the instantiated sizes (A/B runs).
        bool fits = false;
        const int env_w = getenv("GSGP_GEN_WARPS") ? atoi(getenv("GSGP_GEN_WARPS")) : 0;
        if (const char *e = getenv("GSGP_GEN_WAVES")) c->gen_waves = std::max(1, atoi(e));
        for (int w : {env_w, 28, 24, 20, 16, 12}) {
            if (w <= 0) continue;
            c->gen_warps = w;
            fits = gen_smem_bytes(c, c->gen_levels, 32, c->gen_warps) <= (size_t)kMaxSmem;
            if (fits || w == env_w) break;
        }
        c->fused = !(cfg->flags & GSGP_FLAG_KEEP_TREE_SEMANTICS) && !(pl && strcmp(pl, "unfused") == 0) && fits;
    }
    // +256 doubles of slack: the unstaged multi-pair path may read (never use) past the last tile
    CUC(cudaMalloc((void **)&c->dX, sizeof(double) * (row * cfg->nvar + 256)));
    CUC(cudaMalloc((void **)&c->dy, sizeof(double) * row));
    CUC(cudaMalloc((void **)&c->dsym, sizeof(double) * (size_t)c->n_symbols));
    CUC(cudaMemsetAsync(c->dX, 0, sizeof(double) * row * cfg->nvar, c->stream));
    CUC(cudaMemsetAsync(c->dy, 0, sizeof(double) * row, c->stream));
    CUC(cudaMemsetAsync(c->dsym, 0, sizeof(double) * (size_t)c->n_symbols, c->stream));
    {   // store (a): two buffers (cur/new), or one buffer + a block of free cases per row when two do not fit (or on request)
        size_t free_b = 0, total_b = 0;
        CUC(cudaMemGetInfo(&free_b, &total_b));
        const size_t store = sizeof(double) * row * P;
        // everything allocated after the store: tile partials, plan/draw pools, NCCL, the unfused workspace floor
        const size_t slack = ((size_t)3 << 30) + sizeof(double) * 2 * (size_t)P * std::max(c->n_tiles, c->n_tiles_gen);
        const bool fits2 = 2 * store + slack <= free_b;
        c->single_store = c->fused && ((cfg->flags & GSGP_FLAG_SINGLE_BUFFER_STORE) || !fits2);
        c->pitch = (int64_t)row;
        if (c->single_store) {
            const size_t per_tile = sizeof(double) * (size_t)kPass * P;      // bytes of one case tile of every individual
            size_t spare = free_b > store + slack ? free_b - store - slack : 0;
            // a block per launch: half the row is plenty (2-3 launches per generation); leave 40 % of the spare memory free
            int64_t tiles = std::min<int64_t>((int64_t)(spare / 10 * 6 / per_tile), (c->n_tiles_gen + 1) / 2);
            if (fits2) tiles = std::max<int64_t>(1, c->n_tiles_gen / 4);     // on request: a quarter of the store
            if (const char *e = getenv("GSGP_SHIFT_TILES")) tiles = atoi(e);
            if (tiles < 1) { fail(c, "not enough device memory for the semantic store (%.1f GB + one case block)", store / 1e9); return bail(); }
            c->shift_tiles = (int32_t)std::min<int64_t>(tiles, c->n_tiles_gen);
            c->pitch = (int64_t)row + (int64_t)c->shift_tiles * kPass;
        }
    }
    for (int b = 0; b < 2; ++b) {
        if (b == 1 && c->single_store) c->dsem[1] = c->dsem[0] + (size_t)c->shift_tiles * kPass;
        else {
            CUC(cudaMalloc((void **)&c->dsem[b], sizeof(double) * (size_t)c->pitch * P));
            CUC(cudaMemsetAsync(c->dsem[b], 0, sizeof(double) * (size_t)c->pitch * P, c->stream));
        }
        CUC(cudaMalloc((void **)&c->dfit[b], sizeof(double) * (2 * (size_t)P + 1)));   // fit | fit_test | best index: one readback
        c->dfit_test[b] = c->dfit[b] + P;
        CUC(cudaMemsetAsync(c->dfit[b], 0, sizeof(double) * (2 * (size_t)P + 1), c->stream));
    }
    CUC(cudaMalloc((void **)&c->dpartial, sizeof(double) * 2 * (size_t)P * std::max(c->n_tiles, c->n_tiles_gen)));
    CUC(cudaMalloc((void **)&c->dsums, sizeof(double) * 2 * P));
    CUC(cudaMalloc((void **)&c->dsums_all, sizeof(double) * 2 * P * (size_t)cfg->world_size));
    CUC(cudaMalloc((void **)&c->dsums2, sizeof(double) * 2 * P));
    CUC(cudaMalloc((void **)&c->dsums_all2, sizeof(double) * 2 * P * (size_t)cfg->world_size));
    CUC(cudaMalloc((void **)&c->dab, sizeof(double) * 2 * P));
    // plan | program offsets | program descriptors of draw set 0 sit in ONE allocation: a host-replay generation uploads
    // the three with a single copy (gsgp_generation_compiled)
    c->pack_bytes = sizeof(gsgp_plan_entry) * (size_t)P + 2 * sizeof(int32_t) * 2 * (size_t)P;
    CUC(cudaMalloc((void **)&c->dpack, c->pack_bytes));
    CUC(cudaMemsetAsync(c->dpack, 0, c->pack_bytes, c->stream));
    c->dplan = reinterpret_cast<gsgp_plan_entry *>(c->dpack);
    c->dplan_buf[0] = c->dplan; c->dplan_buf[1] = c->dplan;
    if (cfg->rng_mode == GSGP_RNG_DEVICE_PHILOX) {
        CUC(cudaMalloc((void **)&c->dplan_buf[1], sizeof(gsgp_plan_entry) * P));
        CUC(cudaMemsetAsync(c->dplan_buf[1], 0, sizeof(gsgp_plan_entry) * P, c->stream));
    }
    c->dprog_off = reinterpret_cast<int32_t *>(c->dpack + sizeof(gsgp_plan_entry) * (size_t)P);
    c->ds[0].prog_desc = c->dprog_off + 2 * (size_t)P;
    const int n_sets = cfg->rng_mode == GSGP_RNG_DEVICE_PHILOX ? kDrawSets : 1;
    for (int i = 1; i < n_sets; ++i) {
        CUC(cudaMalloc((void **)&c->ds[i].prog_desc, sizeof(int32_t) * 2 * P));
        CUC(cudaMemsetAsync(c->ds[i].prog_desc, 0, sizeof(int32_t) * 2 * P, c->stream));
    }
    CUC(cudaMalloc((void **)&c->dindex_best, sizeof(int32_t)));
    CUC(cudaMemsetAsync(c->dindex_best, 0, sizeof(int32_t), c->stream));

    if (cfg->rng_mode == GSGP_RNG_DEVICE_PHILOX) {
        const int d = cfg->max_depth_creation;
        c->post_stride = (1 << (d + 1));                       // nodes <= 2^(d+1)-1
        c->prog_stride = (1 << d);                             // function nodes <= 2^d-1
        c->tree_stride = 4 * (1 << d);                         // ints  <= 4*2^d-3 (main_cpu.go:609-639)
        // tree offsets inside a draw set's pool are int32 in the plan entries (the ABI's tree0_off / tree1_off)
        if ((int64_t)2 * P * c->tree_stride > INT32_MAX || (int64_t)2 * P * c->prog_stride > INT32_MAX) {
            fail(c, "device RNG mode: population %d x depth %d needs tree pools beyond 2^31 entries", P, d);
            return bail();
        }
        for (int i = 0; i < kDrawSets; ++i) {
            CUC(cudaMalloc((void **)&c->ds[i].tree_pool, sizeof(int32_t) * 2 * (size_t)P * c->tree_stride));
            CUC(c->ds[i].prog_pool.reserve((size_t)2 * P * c->prog_stride));
            CUC(cudaMalloc((void **)&c->ds[i].rec, sizeof(DrawRec) * (size_t)P));
            CUC(cudaMalloc((void **)&c->ds[i].cand, sizeof(int32_t) * 2 * (size_t)P * cfg->tournament_size));
        }
        CUC(cudaMalloc((void **)&c->dpost_pool, sizeof(uint16_t) * 2 * (size_t)P * c->post_stride));
        CUC(cudaMalloc((void **)&c->dcompile_scratch, sizeof(uint32_t) * 2 * (size_t)P * c->post_stride));
        {   // the draw kernel is small and on the critical path of the NEXT generation: let its blocks go first
            int lo = 0, hi = 0;
            CUC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            for (int i = 0; i < kDrawStreams; ++i) CUC(cudaStreamCreateWithPriority(&c->draw_stream[i], cudaStreamNonBlocking, hi));
            c->stream2 = c->draw_stream[0];
            // concurrent draws only when each block keeps its scratch in shared memory (the global scratch pools are one copy)
            c->draw_concurrent = 2 * (size_t)c->post_stride * sizeof(uint32_t) * 32 <= (size_t)200 * 1024;
        }
        for (int i = 0; i < kDrawSets; ++i) {
            CUC(cudaEventCreateWithFlags(&c->ev_draw[i], cudaEventDisableTiming));
            CUC(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
            CUC(cudaEventRecord(c->ev_free[i], c->stream));
        }
        std::vector<int32_t> off(2 * (size_t)P);
        for (int t = 0; t < 2 * P; ++t) off[t] = t * c->prog_stride;
        CUC(cudaMemcpyAsync(c->dprog_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice, c->stream));
        CUC(cudaStreamSynchronize(c->stream));
    }

    // random-tree semantic workspace R: two rows per individual of a chunk
    size_t free_b = 0, total_b = 0;
    CUC(cudaMemGetInfo(&free_b, &total_b));
    const size_t pair = 2 * row * sizeof(double);
    size_t budget = cfg->workspace_bytes > 0 ? (size_t)cfg->workspace_bytes
                                             : std::max<size_t>((size_t)1 << 30, free_b / 4);
    if (cfg->flags & GSGP_FLAG_KEEP_TREE_SEMANTICS) budget = pair * P;
    size_t chunk = std::max<size_t>(1, std::min<size_t>((size_t)P, budget / pair));
    chunk = std::min<size_t>(chunk, 65535);
    if (c->fused) chunk = 1;                                          // R never leaves the SM: no workspace
    if (chunk * pair > free_b) { fail(c, "not enough device memory for the tree-semantic workspace"); return bail(); }
    c->chunk_inds = c->fused ? P : (int32_t)chunk;
    CUC(cudaMalloc((void **)&c->dR, chunk * pair));
    CUC(cudaMemsetAsync(c->dR, 0, chunk * pair, c->stream));

    if (cfg->world_size > 1) {
        std::string err;
        if (!g_nccl.load(err)) { fail(c, "%s", err.c_str()); return bail(); }
        ncclUniqueId id; memcpy(&id, cfg->nccl_id, 128);
        ncclResult_t rc = g_nccl.CommInitRank(&c->comm, cfg->world_size, id, cfg->rank);
        if (rc != ncclSuccess) { fail(c, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc)); return bail(); }
        if (setup_peer_exchange(c)) return bail();
    }
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = c;
    return 0;
}

void gsgp_destroy(gsgp_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->p2p) {                                                      // collective, like the creation: nobody unmaps or frees
        unsigned char b = 1; std::vector<unsigned char> all((size_t)c->cfg.world_size);   // a buffer a peer may still touch
        nccl_gather_bytes(c, &b, all.data(), 1);
        for (int r = 0; r < c->cfg.world_size; ++r) if (c->xpeer_base[r]) cudaIpcCloseMemHandle(c->xpeer_base[r]);
        nccl_gather_bytes(c, &b, all.data(), 1);
    }
    cudaFree(c->xbase);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    for (auto &s : c->prof) for (auto &e : s.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    cudaFree(c->dX); cudaFree(c->dy); cudaFree(c->dsym);
    if (c->single_store) c->dsem[1] = nullptr;                        // one allocation
    for (int b = 0; b < 2; ++b) { cudaFree(c->dsem[b]); cudaFree(c->dfit[b]); }  // dfit_test lives in dfit's allocation
    cudaFree(c->dR); cudaFree(c->dpartial); cudaFree(c->dsums); cudaFree(c->dsums_all);
    cudaFree(c->dsums2); cudaFree(c->dsums_all2); cudaFree(c->dab);
    for (int i = 0; i < kDrawStreams; ++i) if (c->draw_stream[i]) cudaStreamSynchronize(c->draw_stream[i]);
    for (int i = 0; i < kDrawSets; ++i) {
        cudaFree(c->ds[i].tree_pool); if (i > 0) cudaFree(c->ds[i].prog_desc); cudaFree(c->ds[i].rec); cudaFree(c->ds[i].cand);   // set 0's descriptors live in dpack
        if (c->ev_draw[i]) cudaEventDestroy(c->ev_draw[i]);
    }
    for (int i = 0; i < kDrawSets; ++i) if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    for (int i = 0; i < kDrawStreams; ++i) if (c->draw_stream[i]) cudaStreamDestroy(c->draw_stream[i]);
    if (c->dplan_buf[1] != c->dplan_buf[0]) cudaFree(c->dplan_buf[1]);
    cudaFree(c->dpack);                                               // dplan | dprog_off | ds[0].prog_desc
    cudaFree(c->dpost_pool); cudaFree(c->dcompile_scratch); cudaFree(c->dindex_best); cudaFree(c->dhist);
    for (double *p : c->run_sem) cudaFree(p);
    for (gsgp_plan_entry *p : c->run_plan) cudaFree(p);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int gsgp_local_counts(const gsgp_ctx *c, int32_t *nrow_local, int32_t *nrow_test_local)
{
    if (!c) return 1;
    if (nrow_local) *nrow_local = c->L.n_tr;
    if (nrow_test_local) *nrow_test_local = c->L.n_te;
    return 0;
}

int gsgp_exchange_mode(const gsgp_ctx *c)
{
    if (!c || c->cfg.world_size <= 1) return 0;
    return c->p2p ? 2 : 1;
}

int gsgp_store_info(const gsgp_ctx *c, int32_t *buffers, int64_t *buffer_bytes, int64_t *scratch_bytes)
{
    if (!c) return 1;
    if (buffers) *buffers = c->single_store ? 1 : 2;
    if (buffer_bytes) *buffer_bytes = (int64_t)sizeof(double) * c->L.n_pad * c->cfg.population_size;
    if (scratch_bytes) *scratch_bytes = c->single_store ? (int64_t)sizeof(double) * kPass * c->shift_tiles * c->cfg.population_size : 0;
    return 0;
}

int gsgp_set_dataset_shard(gsgp_ctx *c, const double *train_rows, const double *test_rows)
{
    if (!c) return fail(nullptr, "null context");
    CU(c, cudaSetDevice(c->cfg.device));
    const int V = c->cfg.nvar;
    const int32_t n[2] = {c->L.n_tr, c->L.n_te};
    const double *src[2] = {train_rows, test_rows};
    const int32_t dst0[2] = {0, c->L.tr_pad};
    for (int s = 0; s < 2; ++s) {
        if (n[s] == 0) continue;
        if (!src[s]) return fail(c, "null dataset pointer");
        double *tmp = nullptr;
        const size_t bytes = sizeof(double) * (size_t)n[s] * (V + 1);
        CU(c, cudaMalloc((void **)&tmp, bytes));
        CU(c, cudaMemcpyAsync(tmp, src[s], bytes, cudaMemcpyHostToDevice, c->stream));
        transpose_dataset_kernel<<<(n[s] + 255) / 256, 256, 0, c->stream>>>(tmp, n[s], V, c->dX, c->dy, c->L.n_pad, dst0[s]);
        c->launch_count++;
        CU(c, cudaGetLastError());
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(tmp);
    }
    // mean of the train targets for linear scaling (avg_tar main_cpu.go:1078-1083): sequential sum of this rank's
    // rows; with several ranks the local sums are gathered and added in rank order
    {
        double local = 0.0;
        for (int32_t i = 0; i < c->L.n_tr; ++i) local += train_rows[(size_t)i * (V + 1) + V];
        double total = local;
        if (c->cfg.world_size > 1) {
            CU(c, cudaMemcpyAsync(c->dsums, &local, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            ncclResult_t rc = g_nccl.AllGather(c->dsums, c->dsums_all, 1, ncclDouble, c->comm, c->stream);
            if (rc != ncclSuccess) return fail(c, "ncclAllGather failed: %s", g_nccl.GetErrorString(rc));
            std::vector<double> all((size_t)c->cfg.world_size);
            CU(c, cudaMemcpyAsync(all.data(), c->dsums_all, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            total = 0.0;
            for (double v : all) total += v;
        }
        c->tbar = c->cfg.nrow > 0 ? total / (double)c->cfg.nrow : 0.0;
    }
    c->have_dataset = true;
    return 0;
}

int gsgp_set_dataset(gsgp_ctx *c, const double *rowmajor)
{
    if (!c) return fail(nullptr, "null context");
    if (!rowmajor) return fail(c, "null dataset pointer");
    const size_t stride = (size_t)c->cfg.nvar + 1;
    return gsgp_set_dataset_shard(c, rowmajor + (size_t)c->tr_lo * stride,
                                  rowmajor + ((size_t)c->cfg.nrow + c->te_lo) * stride);
}

int gsgp_set_constants(gsgp_ctx *c, const double *sym_val, int32_t n_symbols)
{
    if (!c) return fail(nullptr, "null context");
    if (n_symbols != c->n_symbols) return fail(c, "expected %d symbol values, got %d", c->n_symbols, n_symbols);
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaMemcpyAsync(c->dsym, sym_val, sizeof(double) * (size_t)n_symbols, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int gsgp_seed_semantic(gsgp_ctx *c, int32_t ind, const double *sem)
{
    if (check_ready(c)) return 1;
    if (ind < 0 || ind >= c->cfg.population_size) return fail(c, "individual %d out of range", ind);
    double *row = c->dsem[c->cur] + (size_t)ind * c->pitch;
    if (c->L.n_tr)
        CU(c, cudaMemcpyAsync(row, sem + c->tr_lo, sizeof(double) * (size_t)c->L.n_tr, cudaMemcpyHostToDevice, c->stream));
    if (c->L.n_te)
        CU(c, cudaMemcpyAsync(row + c->L.tr_pad, sem + c->cfg.nrow + c->te_lo, sizeof(double) * (size_t)c->L.n_te,
                              cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->fitness_valid = false;
    return 0;
}

int gsgp_seed_semantic_files(gsgp_ctx *c, int32_t first_ind, int32_t n_files, const char *const *paths)
{
    if (check_ready(c)) return 1;
    if (n_files < 0 || first_ind < 0 || first_ind + (int64_t)n_files > c->cfg.population_size) return fail(c, "individual range out of bounds");
    if (n_files == 0) return 0;
    if (!paths) return fail(c, "null path list");
    const size_t N = (size_t)c->cfg.nrow + (size_t)c->cfg.nrow_test;
    PinnedBuf<double> buf[2];
    cudaEvent_t done[2] = {nullptr, nullptr};
    CU(c, buf[0].reserve(N)); CU(c, buf[1].reserve(N));
    CU(c, cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    int rc = 0;
    for (int32_t i = 0; i < n_files && !rc; ++i) {
        const int b = i & 1;
        if (i >= 2 && cudaEventSynchronize(done[b]) != cudaSuccess) { rc = fail(c, "cudaEventSynchronize failed"); break; }
        const int prc = gsgph_read_semantic(paths[i], (int32_t)N, buf[b].p);      // parses while the previous upload runs
        if (prc) {
            rc = fail(c, "%s: %s", paths[i] ? paths[i] : "(null)",
                      prc == 1 ? "cannot open semantic file" : prc == 2 ? "Cannot parse semantic value" : "Not enough values when reading semantic file");
            break;
        }
        double *row = c->dsem[c->cur] + (size_t)(first_ind + i) * c->pitch;
        cudaError_t e = cudaSuccess;
        if (c->L.n_tr) e = cudaMemcpyAsync(row, buf[b].p + c->tr_lo, sizeof(double) * (size_t)c->L.n_tr, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && c->L.n_te)
            e = cudaMemcpyAsync(row + c->L.tr_pad, buf[b].p + c->cfg.nrow + c->te_lo, sizeof(double) * (size_t)c->L.n_te, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(done[b], c->stream);
        if (e != cudaSuccess) rc = fail(c, "semantic upload failed: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
    c->fitness_valid = false;
    return rc;
}

static int eval_trees_into(gsgp_ctx *c, int32_t n, const int32_t *pool, int32_t pool_len, const int32_t *off,
                           const int32_t *len, double *dout, int64_t pitch)
{
    std::vector<Ins> progs;
    std::vector<int32_t> hoff(n), hlen(n);
    int32_t max_sp = 0;
    if (stage_programs(c, n, pool, pool_len, off, len, progs, hoff.data(), hlen.data(), max_sp)) return 1;
    DevBuf<Ins> dp; DevBuf<int32_t> doff, dlen;
    CU(c, dp.reserve(progs.size())); CU(c, doff.reserve(n)); CU(c, dlen.reserve(n));
    CU(c, cudaMemcpyAsync(dp.p, progs.data(), sizeof(Ins) * progs.size(), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(doff.p, hoff.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(dlen.p, hlen.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    if (launch_interp(c, dp.p, doff.p, dlen.p, 0, n, max_sp, dout, pitch, n)) return 1;
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int gsgp_eval_initial(gsgp_ctx *c, int32_t first_ind, int32_t n, const int32_t *tree_pool, int32_t pool_len,
                      const int32_t *tree_off, const int32_t *tree_len)
{
    if (check_ready(c)) return 1;
    if (n < 0 || first_ind < 0 || first_ind + (int64_t)n > c->cfg.population_size) return fail(c, "individual range out of bounds");
    if (n == 0) return 0;
    for (int32_t t = 0; t < n; ++t) if (tree_len[t] <= 0) return fail(c, "individual %d has no tree", first_ind + t);
    c->fitness_valid = false;
    return eval_trees_into(c, n, tree_pool, pool_len, tree_off, tree_len,
                           c->dsem[c->cur] + (size_t)first_ind * c->pitch, c->pitch);
}

int gsgp_eval_trees(gsgp_ctx *c, int32_t n, const int32_t *tree_pool, int32_t pool_len, const int32_t *tree_off,
                    const int32_t *tree_len, double *out)
{
    if (check_ready(c)) return 1;
    if (n <= 0) return 0;
    DevBuf<double> tmp;
    const size_t row = (size_t)c->L.n_pad;
    const int32_t batch = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)1 << 30) / (row * 8)));
    CU(c, tmp.reserve((size_t)batch * row));
    const int32_t nl = c->L.n_tr + c->L.n_te;
    for (int32_t t0 = 0; t0 < n; t0 += batch) {
        const int32_t nt = std::min(batch, n - t0);
        if (eval_trees_into(c, nt, tree_pool, pool_len, tree_off + t0, tree_len + t0, tmp.p, c->L.n_pad)) return 1;
        if (c->L.n_tr)
            CU(c, cudaMemcpy2DAsync(out + (size_t)t0 * nl, sizeof(double) * nl, tmp.p, sizeof(double) * row,
                                    sizeof(double) * c->L.n_tr, nt, cudaMemcpyDeviceToHost, c->stream));
        if (c->L.n_te)
            CU(c, cudaMemcpy2DAsync(out + (size_t)t0 * nl + c->L.n_tr, sizeof(double) * nl, tmp.p + c->L.tr_pad,
                                    sizeof(double) * row, sizeof(double) * c->L.n_te, nt, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

static int read_fitness(gsgp_ctx *c, double *fit, double *fit_test, int32_t *index_best)
{
    const int P = c->cfg.population_size;
    CU(c, c->hfit.reserve(2 * (size_t)P + 1)); CU(c, c->hint.reserve(4));
    // fit | fit_test | best index (the kernels that find the best individual leave its index behind the two vectors): the
    // usual request is one copy
    if (fit && fit_test) CU(c, cudaMemcpyAsync(c->hfit.p, c->dfit[c->cur], sizeof(double) * (2 * (size_t)P + (index_best ? 1 : 0)), cudaMemcpyDeviceToHost, c->stream));
    else {
        if (fit) CU(c, cudaMemcpyAsync(c->hfit.p, c->dfit[c->cur], sizeof(double) * P, cudaMemcpyDeviceToHost, c->stream));
        if (fit_test) CU(c, cudaMemcpyAsync(c->hfit.p + P, c->dfit_test[c->cur], sizeof(double) * P, cudaMemcpyDeviceToHost, c->stream));
        if (index_best) CU(c, cudaMemcpyAsync(c->hfit.p + 2 * (size_t)P, c->dfit[c->cur] + 2 * (size_t)P, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    c->hint.p[1] = 0;
    if (c->p2p) CU(c, cudaMemcpyAsync(c->hint.p + 1, c->xerr, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->hint.p[1]) return fail(c, "peer exchange timed out: a rank did not deliver its partial sums");
    if (fit) memcpy(fit, c->hfit.p, sizeof(double) * P);
    if (fit_test) memcpy(fit_test, c->hfit.p + P, sizeof(double) * P);
    if (index_best) memcpy(index_best, c->hfit.p + 2 * (size_t)P, sizeof(int32_t));
    return 0;
}

int gsgp_fitness_all(gsgp_ctx *c, double *fit, double *fit_test, int32_t *index_best)
{
    if (check_ready(c)) return 1;
    const int P = c->cfg.population_size;
    const double n_cases = (double)(c->L.n_tr + c->L.n_te);
    for (int32_t k0 = 0; k0 < P; k0 += 65535) {
        const int32_t nk = std::min(65535, P - k0);
        if (launch_gsop(c, GS_MODE_FITNESS_ONLY, k0, nk, 0, c->dsem[c->cur], c->dsem[c->cur], 8.0 * nk * n_cases)) return 1;
    }
    if (launch_finalize(c, GS_MODE_FITNESS_ONLY, c->cur, c->cur, c->n_tiles, c->generation)) return 1;
    if (launch_best(c, c->cur, c->generation, true)) return 1;
    c->fitness_valid = true;
    return read_fitness(c, fit, fit_test, index_best);
}

// gsgp_generation / gsgp_generation_compiled: `programs` == NULL -> the trees are compiled here (staging workers)
static int generation_impl(gsgp_ctx *c, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                           const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off, const int32_t *prog_desc_in,
                           double *fit_out, double *fit_test_out, int32_t *index_best_out, bool trusted_programs = false)
{
    if (check_ready(c)) return 1;
    static const bool trace = getenv("GSGP_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    if (!c->fitness_valid) return fail(c, "population not evaluated (gsgp_fitness_all)");
    if (!plan) return fail(c, "null plan");
    const int P = c->cfg.population_size;
    // staging: the compiled path fills one pinned image of the device's plan | offsets | descriptors block
    gsgp_plan_entry *hplan; int32_t *hoff, *hlen;
    if (programs) {
        CU(c, c->hpack.reserve(c->pack_bytes));
        hplan = reinterpret_cast<gsgp_plan_entry *>(c->hpack.p);
        hoff = reinterpret_cast<int32_t *>(c->hpack.p + sizeof(gsgp_plan_entry) * (size_t)P);
        hlen = hoff + 2 * (size_t)P;
    } else {
        CU(c, c->hplan.reserve(P)); CU(c, c->hoff.reserve(2 * (size_t)P)); CU(c, c->hlen.reserve(2 * (size_t)P));
        hplan = c->hplan.p; hoff = c->hoff.p; hlen = c->hlen.p;
    }
    std::vector<int32_t> &off = c->st_off, &len = c->st_len;
    off.resize(2 * (size_t)P); len.resize(2 * (size_t)P);
    for (int k = 0; k < P; ++k) {
        const gsgp_plan_entry &e = plan[k];
        if (e.op != GSGP_OP_XO && e.op != GSGP_OP_MUT && e.op != GSGP_OP_REPR) return fail(c, "plan[%d]: bad operator %d", k, e.op);
        if (e.p1 < 0 || e.p1 >= P) return fail(c, "plan[%d]: parent index %d out of range", k, e.p1);
        const bool computed = !e.elite && e.op != GSGP_OP_REPR;
        if (computed && e.op == GSGP_OP_XO && (e.p2 < 0 || e.p2 >= P)) return fail(c, "plan[%d]: parent index %d out of range", k, e.p2);
        const bool t0 = computed, t1 = computed && e.op == GSGP_OP_MUT;
        if (t0 && e.tree0_len <= 0) return fail(c, "plan[%d]: missing random tree", k);
        if (t1 && e.tree1_len <= 0) return fail(c, "plan[%d]: missing second random tree", k);
        off[2 * k] = t0 ? e.tree0_off : -1; len[2 * k] = t0 ? e.tree0_len : 0;
        off[2 * k + 1] = t1 ? e.tree1_off : -1; len[2 * k + 1] = t1 ? e.tree1_len : 0;
        hplan[k] = e;
    }
    auto t1 = now();
    if (programs) {
        // The caller staged the programs (gsgph_compile_trees / the pipelined planner): check what the kernels rely on
        // for memory safety -- slot ranges, operand ids, spill levels -- copy into the pinned buffers, upload, launch.
        if (!prog_off || !prog_desc_in || n_instructions < 0) return fail(c, "null program tables");
        const uint32_t n_opd = (uint32_t)(c->n_symbols - 4);
        int32_t max_need = 0;
        CU(c, c->hprog.reserve((size_t)n_instructions + 1));
        CU(c, c->ds[0].prog_pool.reserve((size_t)n_instructions + 1));
        for (int32_t t = 0; t < 2 * P; ++t) {
            if (len[(size_t)t] <= 0) { hoff[t] = 0; hlen[t] = 0; continue; }
            if (!tree_pool || off[(size_t)t] < 0 || (int64_t)off[(size_t)t] + len[(size_t)t] > pool_len) return fail(c, "tree %d: tree exceeds the tree pool", t);
            const int32_t d = prog_desc_in[t], n = prog_desc_len(d), need = prog_desc_need(d), o = prog_off[t];
            if (n < 1 || o < 0 || (int64_t)o + n > n_instructions) return fail(c, "program of tree slot %d outside the program pool", t);
            if (need > kMaxStackLevels) return fail(c, "program of tree slot %d needs %d spill levels > %d", t, need, kMaxStackLevels);
            if (n > len[(size_t)t] / 4 + 1) return fail(c, "program of tree slot %d is longer than its tree", t);
            // programs made by the library's own planner (gsgph_replay_generations) skip the per-instruction look
            if (!trusted_programs)
                for (int32_t i = 0; i < n; ++i) {
                    const uint32_t code = programs[2 * (size_t)(o + i)], ab = programs[2 * (size_t)(o + i) + 1];
                    const uint32_t level = code >> kLevelShift, aop = code & 7u;
                    if (aop > (uint32_t)A_MOV || (code & ~((~0u << kLevelShift) | K_LOADA | K_PUSH | K_POP | 7u)) ||
                        (ab & 0xFFFFu) >= n_opd || (ab >> 16) >= n_opd || level >= (uint32_t)std::max(need, 1))
                        return fail(c, "program of tree slot %d: bad instruction %d", t, i);
                }
            hoff[t] = o; hlen[t] = d;
            max_need = std::max(max_need, need);
        }
        if (n_instructions > 0) memcpy(c->hprog.p, programs, sizeof(Ins) * (size_t)n_instructions);
        c->dcur = 0;
        CU(c, cudaMemcpyAsync(c->dpack, c->hpack.p, c->pack_bytes, cudaMemcpyHostToDevice, c->stream));   // plan | offsets | descriptors
        if (n_instructions > 0)
            CU(c, cudaMemcpyAsync(c->ds[0].prog_pool.p, c->hprog.p, sizeof(Ins) * (size_t)n_instructions, cudaMemcpyHostToDevice, c->stream));
        if (c->fused) {
            const double n_cases = (double)(c->L.n_tr + c->L.n_te);
            if (launch_gen(c, 0, P, 0, c->dsem[c->cur], c->dsem[c->cur ^ 1], gen_bytes(hplan, 0, P, n_cases))) return 1;
            if (finish_generation(c)) return 1;
        } else if (run_generation_kernels(c, hplan, max_need)) return 1;
        c->last_plan.assign(plan, plan + P);
        c->last_pool.assign(tree_pool, tree_pool + std::max(pool_len, 0));
        auto t2 = now();
        const int rc = read_fitness(c, fit_out, fit_test_out, index_best_out);
        if (trace) {
            auto us = [](auto x, auto y) { return std::chrono::duration<double, std::micro>(y - x).count(); };
            fprintf(stderr, "gsgp_generation_compiled: validate plan %.0f us, check+upload+launch %.0f, wait+readback %.0f\n",
                    us(t0, t1), us(t1, t2), us(t2, now()));
        }
        return rc;
    }
    // compile the generation's trees (compile_prefix_tree) on the staging workers, straight into pinned memory.
    // GSGP_REPLAY_CHUNKS > 1 cuts the population into chunks so that staging chunk i+1 overlaps gen_kernel of chunk
    // i; measured slower at config 2 (4 launches: 4 x 0.31 ms vs 0.94 ms, X tiles staged 4 times), so 1 by default.
    int32_t max_sp = 0;
    if (!c->stage_pool) {
        // one process per GPU: leave cores for the other ranks' hosts (world_size processes share the box)
        int nt = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / (2 * std::max(1, c->cfg.world_size))));
        if (const char *e = getenv("GSGP_STAGE_THREADS")) nt = std::max(1, std::min(16, atoi(e)));
        c->stage_pool.reset(new StagePool(nt));
    }
    std::vector<StagePool::Job> &jobs = c->stage_pool->jobs();
    const int32_t nj = (int32_t)jobs.size();
    size_t prog_cap = 2 * (size_t)P + 1;                               // upper bound: one instruction per function node, >= 1
    for (int32_t t = 0; t < 2 * P; ++t) if (len[(size_t)t] > 0) prog_cap += (size_t)len[(size_t)t] / 4;
    CU(c, c->hprog.reserve(prog_cap));
    CU(c, c->ds[0].prog_pool.reserve(prog_cap));
    c->dcur = 0;
    CU(c, cudaMemcpyAsync(c->dplan, c->hplan.p, sizeof(gsgp_plan_entry) * P, cudaMemcpyHostToDevice, c->stream));
    static const int32_t env_chunks = getenv("GSGP_REPLAY_CHUNKS") ? atoi(getenv("GSGP_REPLAY_CHUNKS")) : 1;
    const int32_t n_chunks = (c->fused && !c->single_store && P >= 512) ? std::max(1, std::min(env_chunks, 16)) : 1;
    const double n_cases = (double)(c->L.n_tr + c->L.n_te);
    size_t base = 0;
    double us_stage = 0.0;
    for (int32_t ch = 0; ch < n_chunks; ++ch) {
        auto ts = now();
        const int32_t k0 = (int32_t)((int64_t)P * ch / n_chunks), k1 = (int32_t)((int64_t)P * (ch + 1) / n_chunks);
        const int32_t s0 = 2 * k0, ns = 2 * (k1 - k0);
        const bool caller_only = ns <= 512 || nj == 1;                 // a few hundred trees compile in ~20 us
        for (int32_t j = 0; j < nj; ++j) {
            StagePool::Job &J = jobs[(size_t)j];
            J.pool = tree_pool; J.off = off.data(); J.len = len.data(); J.pool_len = pool_len; J.n_symbols = c->n_symbols;
            J.t0 = s0 + (int32_t)((int64_t)ns * j / nj); J.t1 = s0 + (int32_t)((int64_t)ns * (j + 1) / nj);
            if (caller_only) { J.t0 = j == 0 ? s0 : s0 + ns; J.t1 = s0 + ns; }
        }
        if (caller_only) {
            for (int32_t j = 1; j < nj; ++j) { jobs[(size_t)j].out.clear(); jobs[(size_t)j].bad_slot = -1; jobs[(size_t)j].stack_need = 0; }
            c->stage_pool->run_caller_only();
        } else c->stage_pool->run_all();
        const size_t base0 = base;
        for (const StagePool::Job &J : jobs) {
            if (J.bad_slot >= 0) return fail(c, "tree %d: %s", J.bad_slot, J.err);
            max_sp = std::max(max_sp, J.stack_need);
            if (!J.out.empty()) memcpy(c->hprog.p + base, J.out.data(), sizeof(Ins) * J.out.size());
            for (int32_t t = J.t0; t < J.t1; ++t) {
                const int32_t d = J.desc[(size_t)(t - J.t0)];
                c->hoff.p[t] = d ? (int32_t)(base + (size_t)J.rel_off[(size_t)(t - J.t0)]) : 0;
                c->hlen.p[t] = d;
            }
            base += J.out.size();
        }
        us_stage += std::chrono::duration<double, std::micro>(now() - ts).count();
        if (base > base0)
            CU(c, cudaMemcpyAsync(c->ds[0].prog_pool.p + base0, c->hprog.p + base0, sizeof(Ins) * (base - base0), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->dprog_off + s0, c->hoff.p + s0, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->ds[0].prog_desc + s0, c->hlen.p + s0, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, c->stream));
        if (c->fused &&
            launch_gen(c, k0, k1 - k0, 0, c->dsem[c->cur], c->dsem[c->cur ^ 1], gen_bytes(c->hplan.p, k0, k1 - k0, n_cases))) return 1;
    }
    auto t2 = now(), t3 = t2;
    if (c->fused) { if (finish_generation(c)) return 1; }
    else if (run_generation_kernels(c, c->hplan.p, max_sp)) return 1;
    auto t4 = now();
    c->last_plan.assign(plan, plan + P);
    c->last_pool.assign(tree_pool, tree_pool + std::max(pool_len, 0));
    auto t5 = now();
    const int rc = read_fitness(c, fit_out, fit_test_out, index_best_out);
    if (trace) {
        auto us = [](auto x, auto y) { return std::chrono::duration<double, std::micro>(y - x).count(); };
        fprintf(stderr, "gsgp_generation: validate %.0f us, stage+upload+launch chunks %.0f (staging %.0f), finalize launches %.0f, keep plan %.0f, wait+readback %.0f\n",
                us(t0, t1), us(t1, t2), us_stage, us(t3, t4), us(t4, t5), us(t5, now()));
    }
    return rc;
}

int gsgp_generation(gsgp_ctx *c, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                    double *fit_out, double *fit_test_out, int32_t *index_best_out)
{
    return generation_impl(c, plan, tree_pool, pool_len, nullptr, 0, nullptr, nullptr, fit_out, fit_test_out, index_best_out);
}

int gsgp_generation_compiled(gsgp_ctx *c, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                             const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off,
                             const int32_t *prog_desc, double *fit_out, double *fit_test_out, int32_t *index_best_out)
{
    if (check_ready(c)) return 1;
    if (!programs) return fail(c, "null programs");
    return generation_impl(c, plan, tree_pool, pool_len, programs, n_instructions, prog_off, prog_desc, fit_out, fit_test_out, index_best_out);
}

// gsgp_generation_compiled for programs the library compiled itself (gsgph_replay_generations: the pipelined planner's
// output): slot ranges and sizes are still checked, the 7 000 instructions of a generation are not re-read one by one.
// Not part of the public ABI (include/gsgp_b200.h): a caller's own programs go through gsgp_generation_compiled.
extern "C" int gsgp_internal_generation_trusted(gsgp_ctx *c, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                                                const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off,
                                                const int32_t *prog_desc, double *fit_out, double *fit_test_out, int32_t *index_best_out)
{
    if (check_ready(c)) return 1;
    if (!programs) return fail(c, "null programs");
    return generation_impl(c, plan, tree_pool, pool_len, programs, n_instructions, prog_off, prog_desc, fit_out, fit_test_out, index_best_out, true);
}

int gsgp_run_generations(gsgp_ctx *c, int32_t n)
{
    if (check_ready(c)) return 1;
    if (c->cfg.rng_mode != GSGP_RNG_DEVICE_PHILOX) return fail(c, "gsgp_run_generations needs rng_mode = GSGP_RNG_DEVICE_PHILOX");
    if (!c->fitness_valid) return fail(c, "population not evaluated (gsgp_fitness_all)");
    const int32_t max_sp = std::min(kMaxSmemLevels, c->cfg.max_depth_creation <= 6 ? 2 : 3);   // shared-memory spill levels
    if (ensure_hist(c, c->generation + n)) return 1;
    for (int32_t i = 0; i < n; ++i) {
        const int32_t g = c->generation + 1;
        // draws run kDrawSets-1 generations ahead on high-priority streams: those of g+7 start the moment
        // gen_kernel(g) has released the draw set they overwrite and run under the following generations whenever
        // SMs are free (a gen_kernel block fills an SM's shared memory; a draw block cannot squeeze in beside it).
        // Small populations are paced by the draw kernel's latency, hence several draws in flight.
        for (int32_t d = 0; d < kDrawSets - 1; ++d)                            // first calls / after gsgp_draw_plan
            if (c->drawn[(g + d) % kDrawSets] != g + d && launch_draw(c, g + d)) return 1;
        if (launch_select(c, g)) return 1;
        if (run_generation_kernels(c, nullptr, max_sp)) return 1;              // records ev_free[g % sets] behind the kernels that read the set
        if (launch_draw(c, g + kDrawSets - 1)) return 1;                       // (its set was released by generation g-1)
    }
    c->last_plan.clear();
    return 0;
}

int gsgp_sync(gsgp_ctx *c)
{
    if (!c) return fail(nullptr, "null context");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < kDrawStreams; ++i) if (c->draw_stream[i]) CU(c, cudaStreamSynchronize(c->draw_stream[i]));
    return 0;
}

int gsgp_get_fitness(gsgp_ctx *c, double *fit, double *fit_test, int32_t *index_best)
{
    if (check_ready(c)) return 1;
    return read_fitness(c, fit, fit_test, index_best);
}

static int copy_row_out(gsgp_ctx *c, const double *row, double *out)
{
    if (c->L.n_tr) CU(c, cudaMemcpyAsync(out, row, sizeof(double) * (size_t)c->L.n_tr, cudaMemcpyDeviceToHost, c->stream));
    if (c->L.n_te) CU(c, cudaMemcpyAsync(out + c->L.n_tr, row + c->L.tr_pad, sizeof(double) * (size_t)c->L.n_te, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int gsgp_get_semantic(gsgp_ctx *c, int32_t ind, double *out)
{
    if (check_ready(c)) return 1;
    if (ind < 0 || ind >= c->cfg.population_size) return fail(c, "individual %d out of range", ind);
    return copy_row_out(c, c->dsem[c->cur] + (size_t)ind * c->pitch, out);
}

int gsgp_get_tree_semantic(gsgp_ctx *c, int32_t ind, int32_t which, double *out)
{
    if (check_ready(c)) return 1;
    if (!(c->cfg.flags & GSGP_FLAG_KEEP_TREE_SEMANTICS)) return fail(c, "context created without GSGP_FLAG_KEEP_TREE_SEMANTICS");
    if (ind < 0 || ind >= c->cfg.population_size || which < 0 || which > 1) return fail(c, "bad tree selector");
    return copy_row_out(c, c->dR + (size_t)(2 * ind + which) * c->L.n_pad, out);
}

int gsgp_get_plan(gsgp_ctx *c, gsgp_plan_entry *plan_out)
{
    if (check_ready(c)) return 1;
    const int P = c->cfg.population_size;
    if (!c->last_plan.empty()) { memcpy(plan_out, c->last_plan.data(), sizeof(gsgp_plan_entry) * P); return 0; }
    CU(c, cudaMemcpyAsync(plan_out, c->dplan, sizeof(gsgp_plan_entry) * P, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int gsgp_get_tree(gsgp_ctx *c, int32_t ind, int32_t which, int32_t *out, int32_t cap, int32_t *len)
{
    if (check_ready(c)) return 1;
    if (ind < 0 || ind >= c->cfg.population_size || which < 0 || which > 1) return fail(c, "bad tree selector");
    gsgp_plan_entry e;
    if (!c->last_plan.empty()) e = c->last_plan[ind];
    else {
        CU(c, cudaMemcpyAsync(&e, c->dplan + ind, sizeof e, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    const int32_t off = which ? e.tree1_off : e.tree0_off, n = which ? e.tree1_len : e.tree0_len;
    if (len) *len = n;
    if (n <= 0) return 0;
    if (n > cap) return fail(c, "tree buffer too small (%d > %d)", n, cap);
    if (!c->last_plan.empty()) { memcpy(out, c->last_pool.data() + off, sizeof(int32_t) * n); return 0; }
    CU(c, cudaMemcpyAsync(out, c->ds[c->dcur].tree_pool + off, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int gsgp_get_history(gsgp_ctx *c, int32_t first, int32_t n, double *best_fit, double *best_fit_test, int32_t *best_index)
{
    if (check_ready(c)) return 1;
    if (first < 0 || n < 0 || first + n > c->generation + 1 || first + n > c->hist_cap) return fail(c, "history range out of bounds");
    std::vector<BestRec> h(n);
    CU(c, cudaMemcpyAsync(h.data(), c->dhist + first, sizeof(BestRec) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) {
        if (best_fit) best_fit[i] = h[i].fit;
        if (best_fit_test) best_fit_test[i] = h[i].fit_test;
        if (best_index) best_index[i] = h[i].index;
    }
    return 0;
}

int32_t gsgp_generation_index(const gsgp_ctx *c) { return c ? c->generation : -1; }

int gsgp_get_best_semantic_of(gsgp_ctx *c, int32_t generation, double *out)
{
    if (check_ready(c)) return 1;
    if (!c->keep_run) return fail(c, "context created without GSGP_FLAG_KEEP_RUN_HISTORY");
    if (generation < 0 || generation > c->generation || (size_t)generation / kRunHistBlock >= c->run_sem.size())
        return fail(c, "generation %d is not in the run history", generation);
    return copy_row_out(c, c->run_sem[(size_t)generation / kRunHistBlock] + (size_t)(generation % kRunHistBlock) * c->L.n_pad, out);
}

int gsgp_get_plan_of(gsgp_ctx *c, int32_t generation, gsgp_plan_entry *plan_out)
{
    if (check_ready(c)) return 1;
    if (!c->keep_run) return fail(c, "context created without GSGP_FLAG_KEEP_RUN_HISTORY");
    if (generation < 1 || generation > c->generation || (size_t)generation / kRunHistBlock >= c->run_plan.size())
        return fail(c, "generation %d is not in the run history", generation);
    const size_t P = (size_t)c->cfg.population_size;
    CU(c, cudaMemcpyAsync(plan_out, c->run_plan[(size_t)generation / kRunHistBlock] + (size_t)(generation % kRunHistBlock) * P,
                          sizeof(gsgp_plan_entry) * P, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int gsgp_set_fitness(gsgp_ctx *c, const double *fit, const double *fit_test)
{
    if (check_ready(c)) return 1;
    const int P = c->cfg.population_size;
    c->preselected = -1;                                               // a selection made on the old values is void
    if (fit) CU(c, cudaMemcpyAsync(c->dfit[c->cur], fit, sizeof(double) * P, cudaMemcpyHostToDevice, c->stream));
    if (fit_test) CU(c, cudaMemcpyAsync(c->dfit_test[c->cur], fit_test, sizeof(double) * P, cudaMemcpyHostToDevice, c->stream));
    if (launch_best(c, c->cur, c->generation)) return 1;
    CU(c, cudaStreamSynchronize(c->stream));
    c->fitness_valid = true;
    return 0;
}

int gsgp_draw_plan(gsgp_ctx *c, int32_t generation, gsgp_plan_entry *plan_out)
{
    if (check_ready(c)) return 1;
    if (c->cfg.rng_mode != GSGP_RNG_DEVICE_PHILOX) return fail(c, "gsgp_draw_plan needs rng_mode = GSGP_RNG_DEVICE_PHILOX");
    for (int i = 0; i < kDrawStreams; ++i) CU(c, cudaStreamSynchronize(c->draw_stream[i]));   // pre-drawn generations are discarded
    c->preselected = -1;
    CU(c, cudaEventRecord(c->ev_free[generation % kDrawSets], c->stream));
    if (launch_draw(c, generation)) return 1;
    if (launch_select(c, generation)) return 1;
    for (int i = 0; i < kDrawSets; ++i) c->drawn[i] = -1;              // a later run re-draws its own generations
    c->last_plan.clear();
    CU(c, cudaMemcpyAsync(plan_out, c->dplan, sizeof(gsgp_plan_entry) * c->cfg.population_size, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

void *gsgp_stream(gsgp_ctx *c) { return c ? (void *)c->stream : nullptr; }

int gsgp_profile_enable(gsgp_ctx *c, int32_t on)
{
    if (!c) return 1;
    c->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
    if (on) for (auto &s : c->prof) { s.used = 0; s.launches = 0; s.bytes = 0.0; s.ms = 0.0; }
    if (on) {                                                  // events are created up front, not inside the timed region
        for (int k : {GSGP_K_GSOP, GSGP_K_INTERP}) {
            ProfSlot &s = c->prof[k];
            while (s.ev.size() < 1024) {
                cudaEvent_t a, b;
                if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) break;
                s.ev.push_back({a, b});
            }
        }
    }
    return 0;
}

int gsgp_profile_read(gsgp_ctx *c, int32_t kernel, int64_t *launches, double *total_ms, double *algorithmic_bytes)
{
    if (!c || kernel < 0 || kernel >= GSGP_K_COUNT) return 1;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->stream));
    ProfSlot &s = c->prof[kernel];
    double ms = 0.0;
    for (size_t i = 0; i < s.used; ++i) {
        float t = 0.f;
        CU(c, cudaEventElapsedTime(&t, s.ev[i].first, s.ev[i].second));
        ms += t;
    }
    if (launches) *launches = (int64_t)s.used;
    if (total_ms) *total_ms = ms;
    if (algorithmic_bytes) *algorithmic_bytes = s.bytes;
    return 0;
}

int64_t gsgp_launch_count(const gsgp_ctx *c) { return c ? c->launch_count : 0; }

}  // extern "C"
