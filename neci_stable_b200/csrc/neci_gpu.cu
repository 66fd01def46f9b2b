// C ABI + host-side orchestration of the B200-native FCIQMC engine
// (include/neci_gpu.h).  One engine == one rank == one GPU; the walker list is
// resident in HBM as structure-of-arrays, every phase of an iteration is a
// kernel launch on one stream, and the only host synchronisation per
// iteration is the read-back of the statistics vector (plus the spawn counts
// when more than one rank exchanges spawns over NCCL).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>
#include <unordered_set>
#include <dlfcn.h>
#include <nccl.h>
#include "kernels.cuh"

using namespace ng;

namespace {

// ---- NCCL through dlopen: single-GPU runs need no NCCL at all; inside a torch
// process this resolves to the libnccl torch already loaded ---------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define NG_SYM(field, name) *(void **)(&field) = dlsym(lib, name); if (!field) { err = std::string("missing NCCL symbol ") + name; return false; }
        NG_SYM(GetUniqueId, "ncclGetUniqueId") NG_SYM(CommInitRank, "ncclCommInitRank") NG_SYM(CommDestroy, "ncclCommDestroy")
        NG_SYM(AllGather, "ncclAllGather") NG_SYM(Send, "ncclSend") NG_SYM(Recv, "ncclRecv")
        NG_SYM(GroupStart, "ncclGroupStart") NG_SYM(GroupEnd, "ncclGroupEnd") NG_SYM(GetErrorString, "ncclGetErrorString")
#undef NG_SYM
        return true;
    }
};
NcclApi g_nccl;

template <class T> T *dalloc(size_t n) {
    void *p = nullptr;
    if (n == 0) n = 1;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
    return (T *)p;
}

// scratch device buffers of one call: freed on every return path
struct Scratch {
    std::vector<void *> ptrs; bool ok = true;
    template <class T> T *get(size_t n) { T *p = dalloc<T>(n); if (!p) ok = false; else ptrs.push_back(p); return p; }
    ~Scratch() { for (void *p : ptrs) cudaFree(p); }
};

}  // namespace

struct neci_gpu_engine {
    neci_gpu_config cfg;
    int nw = 1, W = 3;
    Params P;
    WalkerList L;
    SpawnBuf SB;
    K1Queues K;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6];
    std::string err;
    // device-owned tables
    std::vector<void *> owned;
    double *d_partials = nullptr, *d_stats = nullptr, *h_stats = nullptr;
    unsigned int *d_ticket = nullptr;
    long long *h_ctr = nullptr;
    int rows_walk = 0, rows_gen = 0, rows_eval = 0, rows_sing = 0;   // K1 kernels (rows_spawn = their sum + rows_heavy)
    int rows_spawn = 0, rows_heavy = 0, rows_compress = 0, rows_annih = 0, rows_insert = 0, rows_list = 0, rows_trial = 0, rows_tau = 0, rows_total = 0;
    int grid_spawn = 0, grid_generic = 0, grid_spmv = 0;
    u32 stamp = 0;
    bool need_rebuild = false;
    bool pchb_full = false;            // PCHB particle selection FULL-FULL (neci_gpu_set_pchb_particles)
    long long n_launch = 0;            // kernels launched by this engine since init
    long long n_resident = 0;          // length of the list in HBM (slots holding valid determinant data)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    long long ht_cap = 0;
    // semi-stochastic
    long long n_core_local = 0, n_core_total = 0, core_displ = 0;
    long long *d_row_ptr = nullptr; int *d_col = nullptr; double *d_val = nullptr;      // CSR rows as the host passed / k_core_ham built them
    // the same matrix column-blocked for k_determ_spmv_blocked (kernels.cuh): chunk pointers, 16-bit columns, values,
    // the CTAs' shares, one partial sum per (block, row)
    long long *d_bptr = nullptr, *d_spmv_work = nullptr; unsigned short *d_bcol = nullptr; double *d_bval = nullptr, *d_spmv_partial = nullptr;
    int spmv_cb = 0, spmv_nb = 0; long long core_nnz = 0, core_nnz_padded = 0;
    int *d_core_slots = nullptr; double *d_vpart = nullptr, *d_vfull = nullptr, *d_vout = nullptr, *d_core_diag = nullptr;
    std::vector<int> core_sizes, core_displs;
    // staging for AoS transfers; large uploads run chunked on a copy stream (neci_gpu_upload_walkers)
    long long *d_aos = nullptr; size_t aos_cap = 0;
    cudaStream_t copy_stream = nullptr; cudaEvent_t copy_ev[3] = {nullptr, nullptr, nullptr};
    long long h_nlist = 0, upload_chunk = 1ll << 21;
    // multi-rank
    ncclComm_t comm = nullptr;
    unsigned long long *d_cnt_all = nullptr, *h_cnt_all = nullptr;
    // peer-memory exchange (NVLink): this rank's inbox block, the peers' mapped inboxes
    bool p2p = false;
    void *p2p_block = nullptr; size_t p2p_seg_words = 0;
    std::vector<void *> p2p_peer_base;
    PeerBox X;
    unsigned int xseq = 0;
    // NECI_GPU_TIMING=1: device time of the three kernels of the peer-memory exchange, printed at finalize
    bool x_timing = false; cudaEvent_t x_ev[4] = {nullptr, nullptr, nullptr, nullptr}; double x_ms[3] = {0, 0, 0}; long long x_n = 0;

    int fail(const char *fmt, ...) {
        char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf; return 1;
    }
    template <class T> T *upload(const T *h, size_t n) {
        T *d = dalloc<T>(n);
        if (!d) return nullptr;
        owned.push_back(d);
        if (n) cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice);
        return d;
    }
    template <class T> T *alloc(size_t n) { T *d = dalloc<T>(n); if (d) owned.push_back(d); return d; }
};

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return e->fail("%s failed: %s", #call, cudaGetErrorString(_e)); } while (0)
#define NCK(call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) return e->fail("%s failed: %s", #call, g_nccl.GetErrorString(_r)); } while (0)

// dispatch on (words per determinant, system type)
#define NG_DISPATCH(e, BODY)                                                                              \
    do {                                                                                                  \
        const int _key = (e)->nw * 10 + (((e)->cfg.system_type != NECI_SYS_FCIDUMP_PCHB) ? (e)->cfg.system_type :               \
                                         (e)->cfg.t_hphf ? NG_SYS_PCHB_HPHF : (e)->pchb_full ? NG_SYS_PCHB_FULL : NECI_SYS_FCIDUMP_PCHB); \
        switch (_key) {                                                                                   \
            case 11: { constexpr int NW = 1, SYS = NECI_SYS_FCIDUMP_PCHB; BODY; } break;                  \
            case 12: { constexpr int NW = 1, SYS = NECI_SYS_HUBBARD_RS; BODY; } break;                    \
            case 13: { constexpr int NW = 1, SYS = NECI_SYS_HUBBARD_K; BODY; } break;                     \
            case 14: { constexpr int NW = 1, SYS = NG_SYS_PCHB_HPHF; BODY; } break;                       \
            case 15: { constexpr int NW = 1, SYS = NG_SYS_PCHB_FULL; BODY; } break;                       \
            case 21: { constexpr int NW = 2, SYS = NECI_SYS_FCIDUMP_PCHB; BODY; } break;                  \
            case 22: { constexpr int NW = 2, SYS = NECI_SYS_HUBBARD_RS; BODY; } break;                    \
            case 23: { constexpr int NW = 2, SYS = NECI_SYS_HUBBARD_K; BODY; } break;                     \
            case 24: { constexpr int NW = 2, SYS = NG_SYS_PCHB_HPHF; BODY; } break;                       \
            case 25: { constexpr int NW = 2, SYS = NG_SYS_PCHB_FULL; BODY; } break;                       \
            default: return (e)->fail("unsupported (nifd, system_type) = (%d, %d)", (e)->nw - 1, (e)->cfg.system_type); \
        }                                                                                                 \
    } while (0)

static int ensure_aos(neci_gpu_engine *e, size_t words) {
    if (words <= e->aos_cap) return 0;
    if (e->d_aos) cudaFree(e->d_aos);
    e->d_aos = nullptr; e->aos_cap = 0;
    CK(cudaMalloc((void **)&e->d_aos, std::max<size_t>(words, 1) * 8));
    e->aos_cap = words;
    return 0;
}

extern "C" {

const char *neci_gpu_last_error(const neci_gpu_engine *e) { return e ? e->err.c_str() : "null engine"; }

int neci_gpu_init(const neci_gpu_config *cfg, neci_gpu_engine **out) {
    if (!cfg || !out) return 1;
    neci_gpu_engine *e = new neci_gpu_engine();
    *out = e;
    e->cfg = *cfg;
    if (cfg->nifd < 0 || cfg->nifd > 1 || cfg->nbasis > 64 * (cfg->nifd + 1) || cfg->nbasis > NG_MAX_BASIS)
        return e->fail("unsupported bit representation: nbasis=%d nifd=%d (need nbasis <= 64*(nifd+1) <= %d)", cfg->nbasis, cfg->nifd, NG_MAX_BASIS);
    if (cfg->niftot != cfg->nifd + 2) return e->fail("niftot must be nifd + 2 (lenof_sign = 1, flags word)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return e->fail("no CUDA device: the engine has no CPU fallback");
    CK(cudaSetDevice(cfg->device));
    e->nw = cfg->nifd + 1; e->W = cfg->niftot + 1;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    if (const char *c = getenv("NECI_GPU_UPLOAD_CHUNK")) { const long long v = atoll(c); if (v >= 32) e->upload_chunk = v; }
    for (auto &v : e->ev) CK(cudaEventCreate(&v));
    CK(cudaEventCreate(&e->ev_t0)); CK(cudaEventCreate(&e->ev_t1));

    Params &P = e->P; memset(&P, 0, sizeof P);
    P.nel = cfg->nel; P.nbasis = cfg->nbasis; P.nocc_alpha = cfg->nocc_alpha; P.nocc_beta = cfg->nocc_beta;
    P.nranks = cfg->nranks; P.rank = cfg->rank; P.balance_blocks = cfg->balance_blocks; P.system_type = cfg->system_type;
    if (cfg->balance_blocks < 1) return e->fail("balance_blocks must be >= 1");
    if (cfg->nranks < 1 || cfg->nranks > NG_MAX_PUSH_RANKS || cfg->rank < 0 || cfg->rank >= cfg->nranks)
        return e->fail("nranks must be 1..%d and 0 <= rank < nranks (got rank %d of %d)", NG_MAX_PUSH_RANKS, cfg->rank, cfg->nranks);
    P.bb_magic = ~0ull / (u64)cfg->balance_blocks;
    P.t_trunc_initiator = cfg->t_trunc_initiator; P.t_all_real_coeff = cfg->t_all_real_coeff;
    P.t_real_spawn_cutoff = cfg->t_real_spawn_cutoff; P.t_death_before_comms = cfg->t_death_before_comms;
    P.t_init_coherent_rule = cfg->t_init_coherent_rule; P.t_no_brillouin = cfg->t_no_brillouin; P.t_exch = cfg->t_exch;
    P.t_semi_stochastic = cfg->t_semi_stochastic; P.t_core_inits = cfg->t_core_inits;
    P.t_tau_search = cfg->t_tau_search; P.t_consider_par_bias = cfg->t_consider_par_bias; P.t_hphf = cfg->t_hphf;
    if (cfg->t_hphf && cfg->system_type != NECI_SYS_FCIDUMP_PCHB) return e->fail("t_hphf is implemented for FCIDUMP/PCHB systems only");
    if (cfg->t_hphf && cfg->nocc_alpha != cfg->nocc_beta) return e->fail("t_hphf needs Ms = 0");
    P.p_singles = P.p_doubles = P.p_parallel = 1.0;          // lattice models: one excitation class (set_pchb overrides)
    P.initiator_walk_no = cfg->initiator_walk_no; P.real_spawn_cutoff = cfg->real_spawn_cutoff;
    P.occupied_thresh = cfg->occupied_thresh; P.av_mc_excits = cfg->av_mc_excits; P.hii = cfg->hii; P.ecore = cfg->ecore;
    P.seed = cfg->seed;
    P.ref[0] = (u64)cfg->ilut_ref[0]; P.ref[1] = (e->nw > 1) ? (u64)cfg->ilut_ref[1] : 0;
    P.random_orb_index = e->upload(cfg->random_orb_index, cfg->nbasis);
    P.lb_mapping = e->upload(cfg->load_balance_mapping, cfg->balance_blocks);

    const long long M = cfg->max_walkers, Ms = cfg->max_spawned;
    if (M >= (1ll << 31) - 2 || Ms >= (1ll << 31) - 2) return e->fail("max_walkers / max_spawned must be < 2^31");
    WalkerList &L = e->L; memset(&L, 0, sizeof L);
    L.cap = M;
    L.det0 = e->alloc<u64>(M); L.det1 = (e->nw > 1) ? e->alloc<u64>(M) : nullptr;
    L.sgn = e->alloc<double>(M); L.flg = e->alloc<int>(M); L.diagH = e->alloc<double>(M); L.offH = e->alloc<double>(M);
    long long hc = 1024; while (hc < 2 * M) hc <<= 1;
    e->ht_cap = hc; L.ht = e->alloc<u64>(hc); L.ht_mask = (u64)hc - 1;
    L.freeA = e->alloc<int>(M + 1); L.freeB = e->alloc<int>(M + 1);
    L.ctr = e->alloc<long long>(C_COUNT);
    SpawnBuf &SB = e->SB; memset(&SB, 0, sizeof SB);
    SB.W = e->W; SB.seg_cap = Ms / cfg->nranks;
    SB.buf = e->alloc<long long>((size_t)Ms * e->W);
    SB.recv = (cfg->nranks > 1) ? e->alloc<long long>((size_t)Ms * e->W) : SB.buf;
    SB.cnt = e->alloc<unsigned long long>(cfg->nranks);
    long long sc = 1024; while (sc < 2 * Ms) sc <<= 1;
    SB.sht_cap = (u64)sc; SB.sht = e->alloc<u64>(sc);
    SB.ins_idx = e->alloc<int>(Ms);
    if (cfg->t_all_real_coeff) {
        SB.acc_hi = e->alloc<long long>(Ms); SB.acc_lo = e->alloc<long long>(Ms);
        if (!SB.acc_hi || !SB.acc_lo) return e->fail("device allocation failed (merge accumulators)");
        CK(cudaMemset(SB.acc_hi, 0, (size_t)Ms * 8)); CK(cudaMemset(SB.acc_lo, 0, (size_t)Ms * 8));
    }
    if (cfg->nranks > 1) {
        SB.stage_cap = Ms; SB.stage = e->alloc<long long>((size_t)Ms * e->W); SB.stage_cnt = e->alloc<unsigned long long>(1);
        if (!SB.stage || !SB.stage_cnt) return e->fail("device allocation failed (spawn staging list)");
        CK(cudaMemset(SB.stage_cnt, 0, 8));
    }
    SB.heavy_cap = 1 << 16; SB.heavy = e->alloc<long long>(2 * SB.heavy_cap);
    {
        // queues between the K1 kernels: one parent per occupied determinant, one QE / QS entry per spawning attempt.
        // Attempts per iteration = walkers (x AvMCExcits) on this rank, which MaxWalkersPart bounds the way it bounds
        // the list (MemoryFacPart x InitWalkers, fcimc_initialisation.fpp:1650); an overflow is reported, not ignored.
        K1Queues &K = e->K; memset(&K, 0, sizeof K);
        K.qe_cap = std::max<long long>(M, 1 << 20); K.qs_cap = std::max<long long>(M, 1 << 20);
        // the parent list is cut into one segment per CTA of k_walk (sized once the launch shape is known, below)
        const long long par_cap = M + 256ll * NG_MAX_PAR_SEG;
        K.par_d0 = e->alloc<u64>(par_cap); K.par_meta = e->alloc<u32>(par_cap);
        if (e->nw > 1) K.par_d1 = e->alloc<u64>(par_cap);
        K.par_cnt = e->alloc<u32>(NG_MAX_PAR_SEG);
        K.qe = e->alloc<u64>((size_t)K.qe_cap * (e->nw + 2)); K.qs = e->alloc<u64>((size_t)K.qs_cap * (e->nw + 1));
        K.cnt = e->alloc<unsigned long long>(4);
        if (!K.par_d0 || !K.par_meta || !K.par_cnt || !K.qe || !K.qs || !K.cnt || (e->nw > 1 && !K.par_d1))
            return e->fail("device allocation failed (attempt queues, max_walkers=%lld)", M);
        CK(cudaMemset(K.par_cnt, 0, NG_MAX_PAR_SEG * 4));
        CK(cudaMemset(K.cnt, 0, 32));
    }
    if (!L.det0 || !L.sgn || !L.flg || !L.diagH || !L.offH || !L.ht || !L.freeA || !L.freeB || !L.ctr || !SB.buf ||
        !SB.recv || !SB.cnt || !SB.sht || !SB.ins_idx || !SB.heavy || (e->nw > 1 && !L.det1))
        return e->fail("device allocation failed (max_walkers=%lld, max_spawned=%lld)", M, Ms);
    CK(cudaMemset(L.ctr, 0, C_COUNT * 8));
    CK(cudaMemset(L.sgn, 0, (size_t)M * 8));
    CK(cudaMemset(L.flg, 0, (size_t)M * 4));
    CK(cudaMemset(L.ht, 0xFF, (size_t)hc * 8));
    CK(cudaMemset(SB.sht, 0, (size_t)sc * 8));
    CK(cudaMemset(SB.cnt, 0, cfg->nranks * 8));
    e->stamp = 0;

    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device));
    const int nsm = prop.multiProcessorCount;
    e->grid_generic = nsm * 8;
    e->grid_spmv = nsm * NG_SPMV_CTAS;        // persistent CTAs of k_determ_spmv_blocked: one per SM (200 KB of shared memory each)
    e->rows_compress = e->grid_generic; e->rows_annih = e->grid_generic;
    e->rows_insert = e->grid_generic; e->rows_list = e->grid_generic;
    // the K1 kernels run as persistent grids: one CTA per resident slot of every SM
    {
        int w = 0, g = 0, ev = 0, sg = 0;
        NG_DISPATCH(e, {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w, k_walk<NW, SYS>, NG_BLOCK, 0));
            CK(cudaFuncSetAttribute(k_generate<NW, SYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gen_smem_bytes<NW>()));
            CK(cudaFuncSetAttribute(k_generate_heavy<NW, SYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((K1_GEN_BLOCK / 32) * sizeof(GenStage<NW>))));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, k_generate<NW, SYS>, K1_GEN_BLOCK, gen_smem_bytes<NW>()));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ev, k_evaluate<NW, SYS>, NG_BLOCK, 0));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sg, k_singles<NW, SYS>, NG_BLOCK, 0));
        });
        e->rows_walk = std::min(NG_MAX_PAR_SEG, nsm * std::max(1, w)); e->rows_gen = nsm * std::max(1, g); e->rows_heavy = e->rows_gen;
        // every CTA of k_walk sees at most ceil(n_list / grid) + 256 slots: its segment of the parent list
        e->K.par_nseg = e->rows_walk;
        e->K.par_seg_cap = (M + e->rows_walk - 1) / e->rows_walk + 256;
        e->rows_eval = nsm * std::max(1, ev); e->rows_sing = nsm * std::max(1, sg);
        e->rows_spawn = e->rows_walk + e->rows_gen + e->rows_eval + e->rows_sing;
    }
    e->rows_trial = e->grid_generic;
    e->rows_tau = e->grid_generic;            // rows of k_death_magnitude, placed before the trial rows
    e->rows_total = e->rows_spawn + e->rows_heavy + e->rows_compress + e->rows_annih + e->rows_insert + e->rows_list + e->rows_tau + e->rows_trial;
    e->d_partials = e->alloc<double>((size_t)e->rows_total * NECI_ST_COUNT);
    e->d_stats = e->alloc<double>(NECI_ST_COUNT + C_COUNT);
    e->d_ticket = e->alloc<unsigned int>(1);
    CK(cudaMemset(e->d_ticket, 0, 4));
    CK(cudaMemset(e->d_partials, 0, (size_t)e->rows_total * NECI_ST_COUNT * 8));
    CK(cudaMallocHost((void **)&e->h_stats, (NECI_ST_COUNT + C_COUNT) * 8));
    e->h_ctr = reinterpret_cast<long long *>(e->h_stats + NECI_ST_COUNT);
    if (cfg->nranks > 1) {
        e->d_cnt_all = e->alloc<unsigned long long>((size_t)cfg->nranks * cfg->nranks);
        CK(cudaMallocHost((void **)&e->h_cnt_all, (size_t)cfg->nranks * cfg->nranks * 8));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}

int neci_gpu_finalize(neci_gpu_engine *e) {
    if (!e) return 0;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    if (e->x_timing && e->x_n > 1)
        fprintf(stderr, "neci_gpu[rank %d]: peer-memory exchange, mean over %lld iterations: partition+push %.4f ms, wait %.4f ms, gather %.4f ms\n",
                e->cfg.rank, e->x_n - 1, e->x_ms[0] / (e->x_n - 1), e->x_ms[1] / (e->x_n - 1), e->x_ms[2] / (e->x_n - 1));
    for (auto &v : e->x_ev) if (v) cudaEventDestroy(v);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (size_t r = 0; r < e->p2p_peer_base.size(); ++r)
        if ((int)r != e->cfg.rank && e->p2p_peer_base[r]) cudaIpcCloseMemHandle(e->p2p_peer_base[r]);
    if (e->p2p_block) cudaFree(e->p2p_block);
    for (void *p : e->owned) cudaFree(p);
    if (e->d_aos) cudaFree(e->d_aos);
    for (auto &v : e->copy_ev) if (v) cudaEventDestroy(v);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->h_stats) cudaFreeHost(e->h_stats);
    if (e->h_cnt_all) cudaFreeHost(e->h_cnt_all);
    for (auto &v : e->ev) if (v) cudaEventDestroy(v);
    if (e->ev_t0) cudaEventDestroy(e->ev_t0);
    if (e->ev_t1) cudaEventDestroy(e->ev_t1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return 0;
}

int neci_gpu_set_system_fcidump(neci_gpu_engine *e, const double *umat, int64_t n_umat, const double *tmat2d) {
    CK(cudaSetDevice(e->cfg.device));
    if (n_umat >= (1ll << 31)) return e->fail("n_umat must be < 2^31 (32-bit UMatInd arithmetic on the device)");
    e->P.umat = e->upload(umat, (size_t)n_umat);
    e->P.tmat = e->upload(tmat2d, (size_t)e->cfg.nbasis * e->cfg.nbasis);
    {
        // Coulomb / exchange tables for the diagonal element: J(i,j) = UMAT(UMatInd(i,j,i,j)), K(i,j) = UMAT(UMatInd(i,j,j,i))
        const int ns = e->cfg.nbasis / 2;
        auto tri = [](long long a, long long b) { return a > b ? a * (a - 1) / 2 + b : b * (b - 1) / 2 + a; };
        auto ind = [&](int i, int j, int k, int l) { return tri(tri(i, k), tri(j, l)); };
        std::vector<double> J((size_t)ns * ns), K((size_t)ns * ns);
        for (int i = 1; i <= ns; ++i)
            for (int j = 1; j <= ns; ++j) {
                const long long a = ind(i, j, i, j), b = ind(i, j, j, i);
                if (a > n_umat || b > n_umat) return e->fail("UMAT too short for %d spatial orbitals", ns);
                J[(size_t)(i - 1) * ns + j - 1] = umat[a - 1]; K[(size_t)(i - 1) * ns + j - 1] = umat[b - 1];
            }
        e->P.jmat = e->upload(J.data(), J.size()); e->P.kmat = e->upload(K.data(), K.size()); e->P.n_spat_sys = ns;
    }
    if (!e->P.umat || !e->P.tmat || !e->P.jmat || !e->P.kmat) return e->fail("integral upload failed");
    return 0;
}

int neci_gpu_set_pchb(neci_gpu_engine *e, int32_t n_spat, int32_t ij_max, int32_t ab_max, const double *probs,
                      const double *bias, const int32_t *alias, const double *p_exch, const int32_t *tgt_orbs,
                      double p_singles, double p_doubles, double p_parallel, int32_t n_classes,
                      const int32_t *class_of_spinorb) {
    CK(cudaSetDevice(e->cfg.device));
    if (n_classes > NG_MAX_CLASSES || n_classes < 1) return e->fail("n_classes %d out of range (1..%d)", n_classes, NG_MAX_CLASSES);
    Params &P = e->P;
    const size_t n = (size_t)ij_max * 3 * ab_max;
    P.n_spat = n_spat; P.ij_max = ij_max; P.ab_max = ab_max;
    if (n_spat > 0xffff) return e->fail("n_spat %d too large", n_spat);
    {
        // interleave probs / bias / alias / tgtOrbs into one 32-byte entry per (ij, sampler, ab)
        std::vector<PchbEntry> tab(n);
        std::vector<PchbPair> pair((size_t)ij_max);
        for (size_t ij = 0; ij < (size_t)ij_max; ++ij) {
            pair[ij].p_exch = p_exch[ij]; pair[ij].nonempty = 0; pair[ij].pad = 0;
            for (int s = 0; s < 3; ++s) {
                const size_t base = (ij * 3 + s) * ab_max;
                if (alias[base] != 0) pair[ij].nonempty |= 1 << s;
                for (int ab = 0; ab < ab_max; ++ab) {
                    PchbEntry &t = tab[base + ab];
                    t.prob = probs[base + ab]; t.bias = bias[base + ab];
                    t.tgt = (u32)tgt_orbs[2 * ab] | ((u32)tgt_orbs[2 * ab + 1] << 16);
                    const int al = alias[base + ab];            // 1-based; 0 marks an empty sampler
                    if (al >= 1 && al <= ab_max) {
                        t.prob_alias = probs[base + al - 1];
                        t.tgt_alias = (u32)tgt_orbs[2 * (al - 1)] | ((u32)tgt_orbs[2 * (al - 1) + 1] << 16);
                    } else { t.prob_alias = 0.0; t.tgt_alias = 0; }
                }
            }
        }
        P.pchb = e->upload(tab.data(), n);
        P.pchb_pair = e->upload(pair.data(), pair.size());
    }
    P.p_singles = p_singles; P.p_doubles = p_doubles; P.p_parallel = p_parallel; P.n_classes = n_classes;
    {
        const int nA = e->cfg.nocc_alpha, nB = e->cfg.nocc_beta;
        const int par = nA * (nA - 1) / 2 + nB * (nB - 1) / 2, AB = nA * nB;
        P.pgen_pair_par = p_parallel / (double)par;          // IEEE quotients, as pick_biased_elecs forms them per draw
        P.pgen_pair_opp = (1.0 - p_parallel) / (double)AB;
        P.magic_nalpha = (nA > 1) ? (u32)((1ull << 32) / (unsigned)nA) + 1u : 0u;
        P.inv_1m_ps = 1.0 / (1.0 - p_singles);
        P.c_par = (p_parallel > 0.0) ? (double)par / p_parallel : 0.0;
        P.c_opp = (p_parallel < 1.0) ? (double)AB / (1.0 - p_parallel) : 0.0;
        // same-spin pair index -> (n1, n2): n1 = ceil((1 + sqrt(9 + 8 idx)) / 2), n2 = idx + 1 - (n1 - 1)(n1 - 2) / 2
        // (pick_biased_elecs, src/excit_gens_int_weighted.F90:770-790)
        const int nmax = std::max(nA, nB), ntri = std::max(1, nmax * (nmax - 1) / 2);
        if (nmax > 255) return e->fail("more than 255 electrons of one spin are not supported");
        std::vector<unsigned short> tt((size_t)ntri);
        for (int idx = 0; idx < ntri; ++idx) {
            const int n1 = (int)std::ceil((1.0 + std::sqrt(9.0 + 8.0 * (double)idx)) / 2.0);
            const int n2 = idx + 1 - ((n1 - 1) * (n1 - 2)) / 2;
            tt[idx] = (unsigned short)(n1 | (n2 << 8));
        }
        P.tri_tab = e->upload(tt.data(), tt.size());
    }
    std::vector<unsigned char> cls(e->cfg.nbasis);
    std::vector<int> start(n_classes + 1, 0), orbs;
    memset(P.class_mask, 0, sizeof P.class_mask);
    for (int c = 0; c < n_classes; ++c) {
        start[c] = (int)orbs.size();
        for (int o = 1; o <= e->cfg.nbasis; ++o)
            if (class_of_spinorb[o - 1] == c) { orbs.push_back(o); P.class_mask[c][(o - 1) / 64] |= 1ull << ((o - 1) % 64); }
    }
    start[n_classes] = (int)orbs.size();
    for (int o = 0; o < e->cfg.nbasis; ++o) cls[o] = (unsigned char)class_of_spinorb[o];
    P.class_of_spinorb = e->upload(cls.data(), cls.size());
    P.class_start = e->upload(start.data(), start.size());
    P.class_orbs = e->upload(orbs.data(), orbs.size());
    if (!P.pchb || !P.pchb_pair) return e->fail("PCHB table upload failed");
    return 0;
}

int neci_gpu_set_pchb_particles(neci_gpu_engine *e, int32_t mode, const double *p_first, const double *p_second) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->cfg.system_type != NECI_SYS_FCIDUMP_PCHB || !e->P.pchb) return e->fail("set_pchb_particles: call neci_gpu_set_pchb first (FCIDUMP/PCHB systems)");
    if (mode == 0) { e->pchb_full = false; e->P.pchb_particles = 0; return 0; }
    if (mode != 1 && mode != 2) return e->fail("set_pchb_particles: mode %d is not implemented (0 UNIF-UNIF, 1 FULL-FULL, 2 UNIF-FULL)", mode);
    if (e->cfg.t_hphf) return e->fail("set_pchb_particles: weighted particle selection is not available with t_hphf");
    if (!p_first || !p_second) return e->fail("set_pchb_particles: weighted particle selection needs both probability tables");
    e->P.pchb_particles = mode;
    const size_t nb = (size_t)e->cfg.nbasis;
    e->P.pchb_pfirst = e->upload(p_first, nb);
    e->P.pchb_psecond = e->upload(p_second, nb * nb);
    if (!e->P.pchb_pfirst || !e->P.pchb_psecond) return e->fail("set_pchb_particles: table upload failed");
    e->pchb_full = true;
    // the launch shapes of the K1 kernels were taken for the UNIF-UNIF variant: this variant has its own
    int g = 0, ev = 0, sg = 0;
    NG_DISPATCH(e, {
        CK(cudaFuncSetAttribute(k_generate<NW, SYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gen_smem_bytes<NW>()));
        CK(cudaFuncSetAttribute(k_generate_heavy<NW, SYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((K1_GEN_BLOCK / 32) * sizeof(GenStage<NW>))));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, k_generate<NW, SYS>, K1_GEN_BLOCK, gen_smem_bytes<NW>()));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ev, k_evaluate<NW, SYS>, NG_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sg, k_singles<NW, SYS>, NG_BLOCK, 0));
    });
    // grids may only shrink: the rows of the statistics partials were laid out for the original shapes
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, e->cfg.device));
    const int nsm = prop.multiProcessorCount;
    e->rows_gen = std::min(e->rows_gen, nsm * std::max(1, g)); e->rows_heavy = std::min(e->rows_heavy, e->rows_gen);
    e->rows_eval = std::min(e->rows_eval, nsm * std::max(1, ev)); e->rows_sing = std::min(e->rows_sing, nsm * std::max(1, sg));
    CK(cudaMemset(e->d_partials, 0, (size_t)e->rows_total * NECI_ST_COUNT * 8));      // rows that no kernel writes any more
    return 0;
}

int neci_gpu_set_excit_probs(neci_gpu_engine *e, double p_singles, double p_doubles, double p_parallel) {
    if (e->cfg.system_type != NECI_SYS_FCIDUMP_PCHB) return e->fail("set_excit_probs: FCIDUMP/PCHB systems only");
    if (!(p_singles > 0.0 && p_singles < 1.0 && p_doubles > 0.0 && p_parallel >= 0.0 && p_parallel <= 1.0))
        return e->fail("set_excit_probs: probabilities out of range");
    Params &P = e->P;
    P.p_singles = p_singles; P.p_doubles = p_doubles; P.p_parallel = p_parallel;
    const int nA = e->cfg.nocc_alpha, nB = e->cfg.nocc_beta;
    const int par = nA * (nA - 1) / 2 + nB * (nB - 1) / 2, AB = nA * nB;
    P.pgen_pair_par = p_parallel / (double)par;
    P.pgen_pair_opp = (1.0 - p_parallel) / (double)AB;
    P.inv_1m_ps = 1.0 / (1.0 - p_singles);
    P.c_par = (p_parallel > 0.0) ? (double)par / p_parallel : 0.0;
    P.c_opp = (p_parallel < 1.0) ? (double)AB / (1.0 - p_parallel) : 0.0;
    return 0;
}

int neci_gpu_set_system_hubbard_rs(neci_gpu_engine *e, int32_t max_neigh, const int32_t *neighbours,
                                   const double *tmat2d, double uhub) {
    CK(cudaSetDevice(e->cfg.device));
    if (max_neigh > 8) return e->fail("max_neigh > 8 not supported");
    e->P.max_neigh = max_neigh;
    e->P.neighbours = e->upload(neighbours, (size_t)max_neigh * e->cfg.nbasis);
    e->P.tmat = e->upload(tmat2d, (size_t)e->cfg.nbasis * e->cfg.nbasis);
    e->P.uhub = uhub;
    return 0;
}

int neci_gpu_set_system_hubbard_k(neci_gpu_engine *e, int32_t n_k, const int32_t *ksum, const int32_t *kdiff,
                                  const double *eps_k, double u_over_n) {
    CK(cudaSetDevice(e->cfg.device));
    e->P.n_k = n_k;
    e->P.ksum = e->upload(ksum, (size_t)n_k * n_k); e->P.kdiff = e->upload(kdiff, (size_t)n_k * n_k);
    e->P.eps_k = e->upload(eps_k, (size_t)n_k); e->P.u_over_n = u_over_n;
    {
        // partial sums of create_ab_list_hubbard's cumulative list (src/k_space_hubbard.F90:1795-1817): every allowed
        // orbital adds |U/N|, so the k-th partial sum is k repeated additions -- tabulated in that order of operations
        std::vector<double> cum((size_t)e->cfg.nbasis + 1, 0.0);
        const double w = std::fabs(u_over_n);
        for (int k = 1; k <= e->cfg.nbasis; ++k) cum[k] = cum[k - 1] + w;
        e->P.kcum = e->upload(cum.data(), cum.size());
    }
    {
        // the momentum reflection k -> k_pair - k as a byte-wise look-up table on k-point bit masks
        if (n_k > 64) return e->fail("n_k > 64 not supported");
        const int nb = (n_k + 7) / 8;
        std::vector<u64> perm((size_t)n_k * nb * 256, 0ull);
        for (int kij = 0; kij < n_k; ++kij)
            for (int j = 0; j < nb; ++j)
                for (int v = 0; v < 256; ++v) {
                    u64 m = 0;
                    for (int t = 0; t < 8; ++t) {
                        const int k = 8 * j + t;
                        if (((v >> t) & 1) && k < n_k) m |= 1ull << kdiff[(size_t)kij * n_k + k];
                    }
                    perm[((size_t)kij * nb + j) * 256 + v] = m;
                }
        e->P.kperm = e->upload(perm.data(), perm.size());
        e->P.kperm_bytes = nb;
    }
    return 0;
}

// -------------------------------------------------------------------------------
// the list in the AoS staging buffer (n records, optionally followed by the two gdata rows) becomes the resident list
static int take_in_staged_list(neci_gpu_engine *e, int64_t n, double *dgd, double *dgo);
static int take_in_begin(neci_gpu_engine *e, int64_t n);
static int take_in_range(neci_gpu_engine *e, int64_t i0, int64_t i1, double *dgd, double *dgo);
int neci_gpu_upload_walkers(neci_gpu_engine *e, const int64_t *current_dets, int64_t n, const double *gd, const double *go) {
    CK(cudaSetDevice(e->cfg.device));
    if (n > e->cfg.max_walkers - 1) return e->fail("upload of %lld walkers exceeds max_walkers", (long long)n);
    const size_t words = (size_t)n * e->W;
    if (ensure_aos(e, words + 2 * (size_t)n)) return 1;
    double *dgd = gd ? (double *)(e->d_aos + words) : nullptr, *dgo = go ? (double *)(e->d_aos + words + n) : nullptr;
    // Large lists arrive in chunks on a copy stream while the engine's stream clears the hash table and takes in the
    // chunks that have landed (k_upload per chunk): the SoA conversion and the hash inserts hide behind the host link
    const long long chunk = e->upload_chunk;                 // 2M records (48 MB at W = 3) unless NECI_GPU_UPLOAD_CHUNK says otherwise
    if (n <= chunk) {
        CK(cudaMemcpyAsync(e->d_aos, current_dets, words * 8, cudaMemcpyHostToDevice, e->stream));
        if (gd) CK(cudaMemcpyAsync(dgd, gd, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
        if (go) CK(cudaMemcpyAsync(dgo, go, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
        return take_in_staged_list(e, n, dgd, dgo);
    }
    if (!e->copy_stream) {
        CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (auto &v : e->copy_ev) CK(cudaEventCreateWithFlags(&v, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(e->copy_ev[0], e->stream));           // the staging buffer is free once earlier work has finished
    CK(cudaStreamWaitEvent(e->copy_stream, e->copy_ev[0], 0));
    if (take_in_begin(e, n)) return 1;
    int k = 0;
    for (long long i0 = 0; i0 < n; i0 += chunk, ++k) {
        const long long i1 = std::min<long long>(n, i0 + chunk);
        CK(cudaMemcpyAsync(e->d_aos + (size_t)i0 * e->W, current_dets + (size_t)i0 * e->W, (size_t)(i1 - i0) * e->W * 8, cudaMemcpyHostToDevice, e->copy_stream));
        if (gd) CK(cudaMemcpyAsync(dgd + i0, gd + i0, (size_t)(i1 - i0) * 8, cudaMemcpyHostToDevice, e->copy_stream));
        if (go) CK(cudaMemcpyAsync(dgo + i0, go + i0, (size_t)(i1 - i0) * 8, cudaMemcpyHostToDevice, e->copy_stream));
        cudaEvent_t ev = e->copy_ev[1 + (k & 1)];
        CK(cudaEventRecord(ev, e->copy_stream));
        CK(cudaStreamWaitEvent(e->stream, ev, 0));
        if (take_in_range(e, i0, i1, dgd, dgo)) return 1;
    }
    e->n_resident = n;
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
// clears counters and hash table for a list of n records (on the engine's stream)
static int take_in_begin(neci_gpu_engine *e, int64_t n) {
    CK(cudaMemsetAsync(e->L.ctr, 0, C_COUNT * 8, e->stream));
    CK(cudaMemsetAsync(e->L.ht, 0xFF, (size_t)e->ht_cap * 8, e->stream));
    e->h_nlist = n;                                          // pinned: the copy below reads it when it runs
    CK(cudaMemcpyAsync(&e->L.ctr[C_NLIST], &e->h_nlist, 8, cudaMemcpyHostToDevice, e->stream));
    return 0;
}
// slots [i0, i1) of the staged list become resident (SoA conversion, gdata, hash inserts, holes to the FreeSlot stack)
static int take_in_range(neci_gpu_engine *e, int64_t i0, int64_t i1, double *dgd, double *dgo) {
    const int grid = (int)std::min<long long>(e->grid_generic, std::max<long long>(1, (i1 - i0 + 255) / 256));
    e->n_launch += 1;
    NG_DISPATCH(e, (k_upload<NW, SYS><<<grid, 256, 0, e->stream>>>(e->P, e->L, e->d_aos, i0, i1, dgd, dgo, e->W, e->n_resident)));
    CK(cudaGetLastError());
    return 0;
}
static int take_in_staged_list(neci_gpu_engine *e, int64_t n, double *dgd, double *dgo) {
    if (take_in_begin(e, n)) return 1;
    if (take_in_range(e, 0, n, dgd, dgo)) return 1;
    e->n_resident = n;
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int neci_gpu_synthetic_list(neci_gpu_engine *e, int64_t n_dets_total, uint64_t seed, int64_t *n_local_out) {
    CK(cudaSetDevice(e->cfg.device));
    if (n_dets_total < 0) return e->fail("synthetic_list: negative size");
    // this rank's share with 10 % + 4096 of slack for the fluctuation of the hash partition
    const long long cap = std::min<long long>(e->cfg.max_walkers - 1, (long long)(1.1 * (double)n_dets_total / e->cfg.nranks) + 4096);
    if (ensure_aos(e, (size_t)cap * e->W)) return 1;
    unsigned long long *d_count = dalloc<unsigned long long>(1);
    if (!d_count) return e->fail("allocation failed");
    cudaMemsetAsync(d_count, 0, 8, e->stream);
    const int n_spat = e->cfg.nbasis / 2;
    if (e->nw == 1) k_synth_records<1><<<e->grid_generic, 256, 0, e->stream>>>(e->P, seed, n_dets_total, n_spat, e->d_aos, e->W, cap, d_count);
    else k_synth_records<2><<<e->grid_generic, 256, 0, e->stream>>>(e->P, seed, n_dets_total, n_spat, e->d_aos, e->W, cap, d_count);
    unsigned long long cnt = 0;
    cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t rc = cudaStreamSynchronize(e->stream);
    cudaFree(d_count);
    if (rc != cudaSuccess || cudaGetLastError() != cudaSuccess) return e->fail("synthetic_list: generation failed: %s", cudaGetErrorString(rc));
    if ((long long)cnt > cap) return e->fail("synthetic_list: this rank owns %llu determinants, more than max_walkers allows (%lld)", cnt, cap);
    const long long n = (long long)cnt;
    CK(cudaMemsetAsync(e->L.ht, 0xFF, (size_t)e->ht_cap * 8, e->stream));
    if (n > 0) {
        const int grid = (int)std::min<long long>(e->grid_generic, (n + 255) / 256);
        if (e->nw == 1) k_synth_dedupe<1><<<grid, 256, 0, e->stream>>>(e->L, e->d_aos, e->W, n);
        else k_synth_dedupe<2><<<grid, 256, 0, e->stream>>>(e->L, e->d_aos, e->W, n);
        CK(cudaGetLastError());
    }
    e->n_launch += 2;
    e->n_resident = 0;                      // nothing of an earlier list may be reused by k_upload
    if (n_local_out) *n_local_out = n;
    return take_in_staged_list(e, n, nullptr, nullptr);
}

int neci_gpu_download_walkers(neci_gpu_engine *e, int64_t *current_dets, int64_t *n_out, double *gd, double *go) {
    CK(cudaSetDevice(e->cfg.device));
    long long n = 0;
    CK(cudaMemcpyAsync(&n, &e->L.ctr[C_NLIST], 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (n_out) *n_out = n;
    if (current_dets && n > 0) {
        if (ensure_aos(e, (size_t)n * e->W)) return 1;
        const int grid = (int)std::min<long long>(e->grid_generic, (n + 255) / 256);
        e->n_launch += 1;
        if (e->nw == 1) k_download<1><<<grid, 256, 0, e->stream>>>(e->L, e->d_aos, n, e->W);
        else k_download<2><<<grid, 256, 0, e->stream>>>(e->L, e->d_aos, n, e->W);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(current_dets, e->d_aos, (size_t)n * e->W * 8, cudaMemcpyDeviceToHost, e->stream));
    }
    if (gd && n > 0) CK(cudaMemcpyAsync(gd, e->L.diagH, (size_t)n * 8, cudaMemcpyDeviceToHost, e->stream));
    if (go && n > 0) CK(cudaMemcpyAsync(go, e->L.offH, (size_t)n * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int neci_gpu_download_occupied(neci_gpu_engine *e, double min_weight, int64_t *dets_out, int64_t *n_out, double *gd, double *go) {
    CK(cudaSetDevice(e->cfg.device));
    long long n = 0;
    CK(cudaMemcpyAsync(&n, &e->L.ctr[C_NLIST], 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const long long nchunks = std::max<long long>(1, (n + NG_POPS_CHUNK - 1) / NG_POPS_CHUNK);
    Scratch sc;
    int *d_cnt = sc.get<int>((size_t)nchunks); long long *d_tot = sc.get<long long>(1);
    if (!sc.ok) return e->fail("download_occupied: device allocation failed");
    const int grid = (int)std::min<long long>(e->grid_generic, nchunks);
    e->n_launch += 2;
    k_pops_count<<<grid, 256, 0, e->stream>>>(e->L, min_weight, d_cnt);
    k_pops_scan<<<1, 1024, 0, e->stream>>>(e->L, d_cnt, d_tot);
    CK(cudaGetLastError());
    long long tot = 0;
    CK(cudaMemcpyAsync(&tot, d_tot, 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (n_out) *n_out = tot;
    if (dets_out && tot > 0) {
        // staging: records, then the two gdata rows
        if (ensure_aos(e, (size_t)tot * e->W + 2 * (size_t)tot)) return 1;
        double *dgd = gd ? (double *)(e->d_aos + (size_t)tot * e->W) : nullptr;
        double *dgo = go ? (double *)(e->d_aos + (size_t)tot * e->W + tot) : nullptr;
        e->n_launch += 1;
        if (e->nw == 1) k_pops_write<1><<<grid, 256, 0, e->stream>>>(e->L, min_weight, d_cnt, e->d_aos, e->W, dgd, dgo);
        else k_pops_write<2><<<grid, 256, 0, e->stream>>>(e->L, min_weight, d_cnt, e->d_aos, e->W, dgd, dgo);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(dets_out, e->d_aos, (size_t)tot * e->W * 8, cudaMemcpyDeviceToHost, e->stream));
        if (gd) CK(cudaMemcpyAsync(gd, dgd, (size_t)tot * 8, cudaMemcpyDeviceToHost, e->stream));
        if (go) CK(cudaMemcpyAsync(go, dgo, (size_t)tot * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return 0;
}

// -------------------------------------------------------------------------------
// Shared part of the two ways a core space arrives (neci_gpu_set_core_space: rows built by the host;
// neci_gpu_build_core_space: rows built here): sizes/displacements, the replicated core determinants with their hash
// table, the vectors of determ_projection and the slots of this rank's core determinants in the walker list.
static int core_space_layout(neci_gpu_engine *e, const int32_t *sizes, const int32_t *displs) {
    e->core_sizes.assign(sizes, sizes + e->cfg.nranks); e->core_displs.assign(displs, displs + e->cfg.nranks);
    e->n_core_total = 0; for (int r = 0; r < e->cfg.nranks; ++r) e->n_core_total += sizes[r];
    e->n_core_local = sizes[e->cfg.rank];
    e->core_displ = displs[e->cfg.rank];
    return 0;
}
// Column-blocked copy of the core Hamiltonian for k_determ_spmv_blocked (kernels.cuh): called once the CSR rows are
// in HBM.  Set-up code: the prefix sum over the (block, row) chunk lengths and the CTAs' shares are taken on the host.
static int core_space_blocked(neci_gpu_engine *e, long long nnz) {
    const long long n_local = e->n_core_local, n_core = e->n_core_total;
    e->core_nnz = nnz;
    if (n_local == 0) return 0;
    int nb = (int)((n_core + NG_SPMV_CB_MAX - 1) / NG_SPMV_CB_MAX);
    if (nb > NG_SPMV_NB_MAX) return e->fail("core space of %lld determinants needs more than %d column blocks", n_core, NG_SPMV_NB_MAX);
    int cb = (int)((n_core + nb - 1) / nb);
    cb = (cb + 31) & ~31;
    e->spmv_cb = cb; e->spmv_nb = nb;
    const size_t nchunk = (size_t)nb * n_local;
    e->d_bptr = e->alloc<long long>(nchunk + 1);
    e->d_spmv_partial = e->alloc<double>(nchunk);
    e->d_spmv_work = e->alloc<long long>((size_t)e->grid_spmv + 1);
    if (!e->d_bptr || !e->d_spmv_partial || !e->d_spmv_work) return e->fail("no memory for the column-blocked core Hamiltonian (%lld elements)", nnz);
    const int grid = (int)std::max<long long>(1, std::min<long long>(e->grid_generic, (n_local + 7) / 8));
    k_spmv_block_count<<<grid, 256, 0, e->stream>>>(e->d_row_ptr, e->d_col, n_local, cb, nb, e->d_bptr);
    CK(cudaGetLastError());
    std::vector<long long> bp(nchunk + 1, 0);
    CK(cudaMemcpyAsync(bp.data(), e->d_bptr, nchunk * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    // chunks start at multiples of four elements (256-bit loads of the values) and are padded with zeros up to there
    long long run = 0, found = 0;
    for (size_t k = 0; k < nchunk; ++k) { const long long len = bp[k]; bp[k] = run; run += (len + 3) & ~3ll; found += len; }
    bp[nchunk] = run;
    if (found != nnz) return e->fail("column blocking lost elements (%lld of %lld): column index outside the core space?", found, nnz);
    e->core_nnz_padded = run;
    CK(cudaMemcpyAsync(e->d_bptr, bp.data(), (nchunk + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    e->d_bval = e->alloc<double>((size_t)run + 4); e->d_bcol = e->alloc<unsigned short>((size_t)run + 4);
    if (!e->d_bval || !e->d_bcol) return e->fail("no memory for the column-blocked core Hamiltonian (%lld elements)", run);
    {
        // chunk order first (stable scatter of the rows), then the bank-aware order inside every chunk
        Scratch sc;
        double *t_val = sc.get<double>((size_t)run + 4); unsigned short *t_col = sc.get<unsigned short>((size_t)run + 4);
        if (!sc.ok) return e->fail("no memory for the column-blocked core Hamiltonian (%lld elements, set-up copy)", run);
        k_spmv_block_fill<<<grid, 256, 0, e->stream>>>(e->d_row_ptr, e->d_col, e->d_val, n_local, cb, nb, e->d_bptr, t_val, t_col);
        CK(cudaGetLastError());
        const int grid2 = (int)std::max<long long>(1, std::min<long long>(e->grid_generic, ((long long)nchunk + 7) / 8));
        k_spmv_bank_order<<<grid2, 256, 0, e->stream>>>(e->d_bptr, (long long)nchunk, t_val, t_col, e->d_bval, e->d_bcol);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        e->n_launch += 1;
    }
    // CTA p takes the chunks whose first element lies in [nnz p / G, nnz (p + 1) / G): equal shares of the bytes
    const int G = e->grid_spmv;
    std::vector<long long> work((size_t)G + 1);
    for (int p = 0; p <= G; ++p) {
        const long long x = (long long)(((__int128)run * p) / G);
        work[p] = (p == 0) ? 0 : (p == G) ? (long long)nchunk : (long long)(std::lower_bound(bp.begin(), bp.begin() + nchunk, x) - bp.begin());
    }
    CK(cudaMemcpyAsync(e->d_spmv_work, work.data(), work.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaFuncSetAttribute(k_determ_spmv_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, NG_SPMV_CB_MAX * 8));
    CK(cudaStreamSynchronize(e->stream));
    e->n_launch += 2;
    return 0;
}
static int core_space_tables(neci_gpu_engine *e, const int64_t *core_iluts) {
    const int64_t n_local = e->n_core_local;
    e->d_core_slots = e->alloc<int>((size_t)n_local);
    e->d_vpart = e->alloc<double>((size_t)n_local); e->d_vout = e->alloc<double>((size_t)n_local);
    e->d_vfull = e->alloc<double>((size_t)e->n_core_total);
    // the whole core space is replicated on every rank; this rank's determinants are [displ, displ + n_local)
    long long *d_il_all = e->upload((const long long *)core_iluts, (size_t)e->n_core_total * e->nw);
    long long *d_il = d_il_all + (size_t)e->core_displ * e->nw;
    long long hc = 1024; while (hc < 2 * e->n_core_total) hc <<= 1;
    int *d_cht = e->alloc<int>((size_t)hc);
    if (!d_il_all || !d_cht) return e->fail("core-space upload failed");
    CK(cudaMemsetAsync(d_cht, 0, (size_t)hc * 4, e->stream));
    if (e->n_core_total > 0) {
        const int grid = (int)std::min<long long>(e->grid_generic, (e->n_core_total + 255) / 256);
        if (e->nw == 1) k_core_ht_build<1><<<grid, 256, 0, e->stream>>>(d_il_all, e->n_core_total, d_cht, (u64)hc - 1);
        else k_core_ht_build<2><<<grid, 256, 0, e->stream>>>(d_il_all, e->n_core_total, d_cht, (u64)hc - 1);
        CK(cudaGetLastError());
    }
    e->P.core_iluts = d_il_all; e->P.core_ht = d_cht; e->P.core_ht_mask = (u64)hc - 1;
    if (n_local > 0) {
        const int grid = (int)std::min<long long>(e->grid_generic, (n_local + 255) / 256);
        if (e->nw == 1) k_core_locate<1><<<grid, 256, 0, e->stream>>>(e->P, e->L, d_il, n_local, e->d_core_slots);
        else k_core_locate<2><<<grid, 256, 0, e->stream>>>(e->P, e->L, d_il, n_local, e->d_core_slots);
        CK(cudaGetLastError());
    }
    long long errf = 0;
    CK(cudaMemcpyAsync(&errf, &e->L.ctr[C_ERR], 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (errf & 64) return e->fail("core determinant missing from the uploaded walker list");
    return 0;
}
int neci_gpu_set_core_space(neci_gpu_engine *e, int64_t n_local, const int64_t *row_ptr, const int32_t *col,
                            const double *val, const int32_t *sizes, const int32_t *displs, const int64_t *core_iluts) {
    CK(cudaSetDevice(e->cfg.device));
    core_space_layout(e, sizes, displs);
    if (n_local != e->n_core_local) return e->fail("set_core_space: n_local = %lld but sizes[rank] = %lld", (long long)n_local, (long long)e->n_core_local);
    const long long nnz = row_ptr[n_local];
    e->d_row_ptr = e->upload((const long long *)row_ptr, (size_t)n_local + 1);
    e->d_col = e->alloc<int>((size_t)nnz + 4); e->d_val = e->alloc<double>((size_t)nnz + 4);
    if (!e->d_col || !e->d_val) return e->fail("set_core_space: no memory for %lld non-zero elements", nnz);
    CK(cudaMemset(e->d_col + nnz, 0, 16)); CK(cudaMemset(e->d_val + nnz, 0, 32));
    CK(cudaMemcpy(e->d_col, col, (size_t)nnz * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(e->d_val, val, (size_t)nnz * 8, cudaMemcpyHostToDevice));
    {
        // core_ham_diag (fast_determ_hamil.F90:1494-1507): the diagonal entry of every local row
        std::vector<double> diag((size_t)n_local, 0.0);
        for (int64_t i = 0; i < n_local; ++i)
            for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
                if (col[k] == i + displs[e->cfg.rank]) diag[i] = val[k];
        e->d_core_diag = e->upload(diag.data(), diag.size());
    }
    if (core_space_tables(e, core_iluts)) return 1;
    return core_space_blocked(e, nnz);
}

int neci_gpu_build_core_space(neci_gpu_engine *e, const int32_t *sizes, const int32_t *displs, const int64_t *core_iluts,
                              int64_t *nnz_out) {
    CK(cudaSetDevice(e->cfg.device));
    core_space_layout(e, sizes, displs);
    if (e->n_core_total >= (1ll << 31)) return e->fail("build_core_space: core space too large for int32 columns");
    if (core_space_tables(e, core_iluts)) return 1;
    const long long n_local = e->n_core_local, n_core = e->n_core_total;
    e->d_row_ptr = e->alloc<long long>((size_t)n_local + 1);
    e->d_core_diag = e->alloc<double>((size_t)n_local);
    if (!e->d_row_ptr || !e->d_core_diag) return e->fail("build_core_space: allocation failed");
    const long long *d_il = e->P.core_iluts;
    const double hii = e->cfg.hii;
    const int grid = (int)std::max<long long>(1, std::min<long long>(e->grid_generic, (n_local * 32 + NG_BLOCK - 1) / NG_BLOCK));
    // pass 1: row lengths; the prefix sum over n_local numbers is taken on the host (set-up code, 8 bytes per row)
    NG_DISPATCH(e, (k_core_ham<NW, SYS, false><<<grid, NG_BLOCK, 0, e->stream>>>(e->P, d_il, n_core, e->core_displ, n_local, hii,
                                                                                e->d_row_ptr, nullptr, nullptr, nullptr)));
    CK(cudaGetLastError());
    std::vector<long long> rp((size_t)n_local + 1, 0);
    CK(cudaMemcpyAsync(rp.data(), e->d_row_ptr, (size_t)n_local * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    long long run = 0;
    for (long long i = 0; i < n_local; ++i) { const long long len = rp[i]; rp[i] = run; run += len; }
    rp[n_local] = run;
    CK(cudaMemcpyAsync(e->d_row_ptr, rp.data(), (size_t)(n_local + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    e->d_col = e->alloc<int>((size_t)run + 4); e->d_val = e->alloc<double>((size_t)run + 4);
    if (!e->d_col || !e->d_val) return e->fail("build_core_space: no memory for %lld non-zero elements", run);
    CK(cudaMemsetAsync(e->d_col + run, 0, 16, e->stream)); CK(cudaMemsetAsync(e->d_val + run, 0, 32, e->stream));
    // pass 2: the elements again, written in place
    NG_DISPATCH(e, (k_core_ham<NW, SYS, true><<<grid, NG_BLOCK, 0, e->stream>>>(e->P, d_il, n_core, e->core_displ, n_local, hii,
                                                                               e->d_row_ptr, e->d_col, e->d_val, e->d_core_diag)));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    e->n_launch += 2;
    if (nnz_out) *nnz_out = run;
    return core_space_blocked(e, run);
}

int neci_gpu_get_core_hamiltonian(neci_gpu_engine *e, int64_t *row_ptr, int32_t *col, double *val) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->d_row_ptr) return e->fail("get_core_hamiltonian: no core space set");
    const long long n_local = e->n_core_local;
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(row_ptr, e->d_row_ptr, (size_t)(n_local + 1) * 8, cudaMemcpyDeviceToHost));
    const long long nnz = row_ptr[n_local];
    if (col) CK(cudaMemcpy(col, e->d_col, (size_t)nnz * 4, cudaMemcpyDeviceToHost));
    if (val) CK(cudaMemcpy(val, e->d_val, (size_t)nnz * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int neci_gpu_set_trial_space(neci_gpu_engine *e, int64_t n_trial, const int64_t *trial_iluts, const double *trial_amps,
                             int64_t n_con, const int64_t *con_iluts, const double *con_amps) {
    CK(cudaSetDevice(e->cfg.device));
    if (n_trial < 0 || n_con < 0 || n_trial + n_con >= (1ll << 31) - 2) return e->fail("trial / connected space size out of range");
    const int nw = e->nw;
    // one table for both spaces, trial entries first; a connected determinant that is also a trial determinant is
    // dropped (hash_search_trial looks in the trial table first, src/searching.F90:198-221)
    struct KH { size_t operator()(const std::pair<uint64_t, uint64_t> &k) const { return (size_t)mix64(k.first ^ mix64(k.second + 0x9E3779B97F4A7C15ull)); } };
    std::unordered_set<std::pair<uint64_t, uint64_t>, KH> seen;
    std::vector<long long> il; std::vector<double> amp;
    il.reserve((size_t)(n_trial + n_con) * nw); amp.reserve((size_t)(n_trial + n_con));
    long long nt = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t n = pass ? n_con : n_trial;
        const int64_t *src = pass ? con_iluts : trial_iluts;
        const double *a = pass ? con_amps : trial_amps;
        for (int64_t i = 0; i < n; ++i) {
            const std::pair<uint64_t, uint64_t> key((uint64_t)src[i * nw], nw > 1 ? (uint64_t)src[i * nw + 1] : 0ull);
            if (!seen.insert(key).second) continue;
            for (int w = 0; w < nw; ++w) il.push_back(src[i * nw + w]);
            amp.push_back(a[i]);
        }
        if (!pass) nt = (long long)amp.size();
    }
    const long long n_all = (long long)amp.size();
    long long hc = 1024; while (hc < 2 * n_all) hc <<= 1;
    long long *d_il = e->upload(il.data(), il.size());
    double *d_amp = e->upload(amp.data(), amp.size());
    int *d_ht = e->alloc<int>((size_t)hc);
    if (!e->L.trial_amp) e->L.trial_amp = e->alloc<double>((size_t)e->cfg.max_walkers);
    if (!d_il || !d_amp || !d_ht || !e->L.trial_amp) return e->fail("trial-space upload failed");
    CK(cudaMemsetAsync(d_ht, 0, (size_t)hc * 4, e->stream));
    CK(cudaMemsetAsync(e->L.trial_amp, 0, (size_t)e->cfg.max_walkers * 8, e->stream));
    if (n_all > 0) {
        const int grid = (int)std::min<long long>(e->grid_generic, (n_all + 255) / 256);
        if (nw == 1) k_trial_ht_build<1><<<grid, 256, 0, e->stream>>>(d_il, n_all, d_ht, (u64)hc - 1);
        else k_trial_ht_build<2><<<grid, 256, 0, e->stream>>>(d_il, n_all, d_ht, (u64)hc - 1);
        CK(cudaGetLastError());
    }
    e->P.trial_iluts = d_il; e->P.trial_amps = d_amp; e->P.trial_ht = d_ht; e->P.trial_ht_mask = (u64)hc - 1; e->P.n_trial = nt;
    if (nw == 1) k_trial_locate<1><<<e->grid_generic, 256, 0, e->stream>>>(e->P, e->L);
    else k_trial_locate<2><<<e->grid_generic, 256, 0, e->stream>>>(e->P, e->L);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- spawn exchange: SendProcNewParts (Annihilation.F90:150-247) over NCCL ------
static int exchange_spawns(neci_gpu_engine *e, long long *n_recv_out) {
    const int nr = e->cfg.nranks;
    // MPI_Alltoall of the counts == all-gather of every rank's count vector
    NCK(g_nccl.AllGather(e->SB.cnt, e->d_cnt_all, (size_t)nr, ncclUint64, e->comm, e->stream));
    CK(cudaMemcpyAsync(e->h_cnt_all, e->d_cnt_all, (size_t)nr * nr * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const int me = e->cfg.rank;
    long long total = 0;
    for (int s = 0; s < nr; ++s) total += (long long)std::min<unsigned long long>(e->h_cnt_all[(size_t)s * nr + me], (unsigned long long)e->SB.seg_cap);
    if (total > e->cfg.max_spawned) return e->fail("received spawns exceed max_spawned");
    NCK(g_nccl.GroupStart());
    long long off = 0;
    for (int s = 0; s < nr; ++s) {       // receive buffer ordered by source rank (MPI_Alltoallv displacements)
        const long long cs = (long long)std::min<unsigned long long>(e->h_cnt_all[(size_t)s * nr + me], (unsigned long long)e->SB.seg_cap);
        const long long cd = (long long)std::min<unsigned long long>(e->h_cnt_all[(size_t)me * nr + s], (unsigned long long)e->SB.seg_cap);
        if (cd > 0) NCK(g_nccl.Send(e->SB.buf + (size_t)s * e->SB.seg_cap * e->W, (size_t)cd * e->W, ncclInt64, s, e->comm, e->stream));
        if (cs > 0) NCK(g_nccl.Recv(e->SB.recv + (size_t)off * e->W, (size_t)cs * e->W, ncclInt64, s, e->comm, e->stream));
        off += cs;
    }
    NCK(g_nccl.GroupEnd());
    *n_recv_out = total;
    return 0;
}

// the same exchange over NVLink peer memory: push kernel + mailbox wait + local compaction, no host round trip
static int exchange_spawns_p2p(neci_gpu_engine *e, bool from_stage) {
    const int nr = e->cfg.nranks;
    e->xseq += 1;
    e->n_launch += from_stage ? 2 : 3;
    if (e->x_timing && e->x_n > 0) {                       // the previous exchange has long finished
        float ms;
        for (int k = 0; k < 3; ++k) if (cudaEventElapsedTime(&ms, e->x_ev[k], e->x_ev[k + 1]) == cudaSuccess) e->x_ms[k] += ms;
    }
    if (e->x_timing) cudaEventRecord(e->x_ev[0], e->stream);
    // spawning pass: the spawning kernels have routed and pushed their spawns already (spawn_stage_push); block moves
    // arrive packed per destination in SpawnedParts and are copied into the inboxes here
    if (!from_stage) k_push<<<e->grid_generic, 256, 0, e->stream>>>(e->SB, e->X, nr, e->cfg.rank, e->xseq);
    if (e->x_timing) cudaEventRecord(e->x_ev[1], e->stream);
    k_wait<<<1, 64, 0, e->stream>>>(e->L, e->SB, e->X, nr, e->cfg.rank, e->xseq, from_stage, 60000000000ll /* ~30 s of SM clocks: ranks may be skewed by I/O */);
    if (e->x_timing) cudaEventRecord(e->x_ev[2], e->stream);
    k_gather<<<e->grid_generic, 256, 0, e->stream>>>(e->SB, e->X, nr, e->xseq);
    if (e->x_timing) { cudaEventRecord(e->x_ev[3], e->stream); e->x_n += 1; }
    CK(cudaGetLastError());
    return 0;
}
// spawn exchange of one iteration: leaves the received records contiguous in SB.recv and the count in A.n_recv
// (>= 0: known on the host; -2: on the device)
static int exchange(neci_gpu_engine *e, long long *n_recv, bool from_stage) {
    if (e->p2p) { *n_recv = -2; return exchange_spawns_p2p(e, from_stage); }
    if (!e->comm) return e->fail("nranks > 1 but neither neci_gpu_nccl_init nor neci_gpu_p2p_open was called");
    if (from_stage) {                       // NCCL needs SpawnedParts' per-destination segments: route locally first
        e->n_launch += 1;
        if (e->nw == 1) k_partition<1><<<e->grid_generic, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB);
        else k_partition<2><<<e->grid_generic, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB);
        CK(cudaGetLastError());
    }
    return exchange_spawns(e, n_recv);
}

static int gather_core_vector(neci_gpu_engine *e) {
    const int nr = e->cfg.nranks;
    if (nr == 1) {
        CK(cudaMemcpyAsync(e->d_vfull, e->d_vpart, (size_t)e->n_core_local * 8, cudaMemcpyDeviceToDevice, e->stream));
        return 0;
    }
    // MPIAllGatherV (semi_stoch_procs.F90:127) as grouped send/recv of the ragged shares
    if (!e->comm) return e->fail("semi-stochastic run on several ranks needs neci_gpu_nccl_init (the core vector is gathered with NCCL)");
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < nr; ++r) {
        if (e->n_core_local > 0) NCK(g_nccl.Send(e->d_vpart, (size_t)e->n_core_local, ncclFloat64, r, e->comm, e->stream));
        if (e->core_sizes[r] > 0) NCK(g_nccl.Recv(e->d_vfull + e->core_displs[r], (size_t)e->core_sizes[r], ncclFloat64, r, e->comm, e->stream));
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}

static int annihilation_phase(neci_gpu_engine *e, IterArgs &A, int row0) {
    const int g = e->grid_generic;
    double *p_comp = e->d_partials + (size_t)row0 * NECI_ST_COUNT;
    double *p_ann = p_comp + (size_t)e->rows_compress * NECI_ST_COUNT;
    double *p_ins = p_ann + (size_t)e->rows_annih * NECI_ST_COUNT;
    double *p_lst = p_ins + (size_t)e->rows_insert * NECI_ST_COUNT;
    e->stamp += 1;
    if ((e->stamp & 0xFFFFu) == 0) { e->stamp += 1; CK(cudaMemsetAsync(e->SB.sht, 0, (size_t)e->SB.sht_cap * 8, e->stream)); }
    A.stamp = e->stamp;
    e->n_launch += 4 + ((e->cfg.t_semi_stochastic && e->n_core_local > 0) ? 1 : 0);
    if (e->nw == 1) k_compress<1><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, A, p_comp, e->d_ticket);
    else k_compress<2><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, A, p_comp, e->d_ticket);
    if (e->cfg.t_semi_stochastic && e->n_core_local > 0)
        k_determ_apply<<<std::max(1, (int)std::min<long long>(g, (e->n_core_local + 255) / 256)), 256, 0, e->stream>>>(e->L, e->d_core_slots, e->d_vout, e->n_core_local);
    if (e->nw == 1) k_annihilate<1><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, A, p_ann);
    else k_annihilate<2><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, A, p_ann);
    NG_DISPATCH(e, (k_insert<NW, SYS><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, p_ins)));
    if (e->nw == 1) k_list_stats<1><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, A, p_lst);
    else k_list_stats<2><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, A, p_lst);
    CK(cudaGetLastError());
    return 0;
}

static int finish_iteration(neci_gpu_engine *e, double *stats_out) {
    e->n_launch += 1;
    if (e->nw == 1) k_reduce_stats<1><<<NECI_ST_COUNT, 256, 0, e->stream>>>(e->P, e->L, e->SB, e->d_partials, e->rows_total, e->d_stats);
    else k_reduce_stats<2><<<NECI_ST_COUNT, 256, 0, e->stream>>>(e->P, e->L, e->SB, e->d_partials, e->rows_total, e->d_stats);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev[4], e->stream));
    // statistics and device counters in one copy (k_reduce_stats leaves the counters behind the statistics)
    CK(cudaMemcpyAsync(e->h_stats, e->d_stats, (NECI_ST_COUNT + C_COUNT) * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    float ms;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->h_stats[NECI_ST_TIME_DETERM_MS] = ms;
    cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]); e->h_stats[NECI_ST_TIME_SPAWN_MS] = ms;
    cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->h_stats[NECI_ST_TIME_COMM_MS] = ms;
    cudaEventElapsedTime(&ms, e->ev[3], e->ev[4]); e->h_stats[NECI_ST_TIME_ANNIHIL_MS] = ms;
    if (stats_out) memcpy(stats_out, e->h_stats, NECI_ST_COUNT * 8);
    e->n_resident = std::min<long long>(e->h_ctr[C_NLIST], e->cfg.max_walkers - 1);
    // tombstones are recycled by inserts; rebuild the table when they pile up
    if (e->h_ctr[C_NTOMB] > e->ht_cap / 4) e->need_rebuild = true;
    const long long errf = e->h_ctr[C_ERR];
    if (errf & 1) return e->fail("spawned-list overflow (increase max_spawned / MemoryFacSpawn)");
    if (errf & 2) return e->fail("main walker list overflow (increase max_walkers / MemoryFacPart)");
    if (errf & 4) return e->fail("death probability > 2: algorithm unstable, reduce tau");
    if (errf & 16) return e->fail("excitation generator could not find an excitation after 250 attempts");
    if (errf & 32) return e->fail("heavy-determinant queue overflow");
    if (errf & 128) return e->fail("spawning-attempt queue overflow (more walkers than max_walkers: increase MemoryFacPart)");
    if (errf & 256) return e->fail("peer-memory spawn exchange timed out waiting for another rank");
    if (errf & 512) return e->fail("a determinant holds 2^29 or more walkers (attempt index field of the spawning queues)");
    return 0;
}

static int begin_iteration(neci_gpu_engine *e) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->need_rebuild) {
        CK(cudaMemsetAsync(e->L.ht, 0xFF, (size_t)e->ht_cap * 8, e->stream));
        long long z = 0;
        CK(cudaMemcpyAsync(&e->L.ctr[C_NTOMB], &z, 8, cudaMemcpyHostToDevice, e->stream));
        if (e->nw == 1) k_ht_rebuild<1><<<e->grid_generic, 256, 0, e->stream>>>(e->P, e->L);
        else k_ht_rebuild<2><<<e->grid_generic, 256, 0, e->stream>>>(e->P, e->L);
        e->need_rebuild = false;
        e->n_launch += 1;
    }
    // ValidSpawnedList = InitialSpawnedSlots etc. (FciMCPar.F90:1237-1248): one launch for all per-iteration counters
    if (e->SB.push_cnt)        // this rank's segment in its peers' inboxes for the exchange that follows the spawning pass
        e->SB.push_off = ((long long)((e->xseq + 1) & 1u) * e->cfg.nranks + e->cfg.rank) * e->SB.seg_cap;
    k_begin_iteration<<<1, 64, 0, e->stream>>>(e->L, e->SB, e->K, e->cfg.nranks);
    e->n_launch += 1;
    return 0;
}

int neci_gpu_iterate(neci_gpu_engine *e, double tau, double diag_sft, int64_t iter, double *stats_out) {
    if (begin_iteration(e)) return 1;
    IterArgs A; A.tau = tau; A.diag_sft = diag_sft; A.iter = iter; A.n_recv = -1; A.stamp = 0;
    CK(cudaEventRecord(e->ev[0], e->stream));
    if (e->cfg.t_semi_stochastic && e->n_core_total > 0) {
        // determ_projection (semi_stoch_procs.F90:105-241): gather of partial_determ_vecs, MPIAllGatherV, multiplication.
        // The phase time between ev[0] and ev[1] is exactly that routine.
        e->n_launch += (e->n_core_local > 0) ? 3 : 0;
        const bool single = e->cfg.nranks == 1;
        if (e->n_core_local > 0)
            k_core_gather<<<std::max(1, (int)std::min<long long>(e->grid_generic, (e->n_core_local + 255) / 256)), 256, 0, e->stream>>>(
                e->L, e->d_core_slots, e->n_core_local, single ? e->d_vfull : e->d_vpart);
        if (!single && gather_core_vector(e)) return 1;
        if (e->n_core_local > 0) {
            const int cols = (int)std::min<long long>(e->spmv_cb, e->n_core_total);
            k_determ_spmv_blocked<<<e->grid_spmv, NG_SPMV_THREADS, (size_t)cols * 8, e->stream>>>(
                e->d_bptr, e->d_bcol, e->d_bval, e->d_vfull, e->d_spmv_work, e->n_core_local, e->n_core_total, e->spmv_cb, e->d_spmv_partial);
            k_determ_finish<<<std::max(1, (int)std::min<long long>(e->grid_generic, (e->n_core_local + 255) / 256)), 256, 0, e->stream>>>(
                e->d_spmv_partial, e->d_vfull, e->n_core_local, e->spmv_nb, e->core_displ, tau, diag_sft,
                e->cfg.t_death_before_comms ? (const double *)nullptr : (const double *)e->d_core_diag, e->d_vout);
        }
    }
    CK(cudaEventRecord(e->ev[1], e->stream));
    if (e->cfg.t_tau_search) {
        double *p_tau = e->d_partials + (size_t)(e->rows_total - e->rows_trial - e->rows_tau) * NECI_ST_COUNT;
        e->n_launch += 1;
        k_death_magnitude<<<e->rows_tau, NG_BLOCK, 0, e->stream>>>(e->L, diag_sft, p_tau);
    }
    if (e->P.trial_ht) {
        // trial part of SumEContrib on the signs the walker loop sees (before death)
        double *p_trial = e->d_partials + (size_t)(e->rows_total - e->rows_trial) * NECI_ST_COUNT;
        e->n_launch += 1;
        if (e->nw == 1) k_trial_energy<1><<<e->rows_trial, NG_BLOCK, 0, e->stream>>>(e->P, e->L, p_trial);
        else k_trial_energy<2><<<e->rows_trial, NG_BLOCK, 0, e->stream>>>(e->P, e->L, p_trial);
    }
    {
        // the loop over determinants as four streaming kernels (spawn_kernel.cuh); PCHB singles have their own
        const bool pchb = e->cfg.system_type == NECI_SYS_FCIDUMP_PCHB;
        e->n_launch += pchb ? 5 : 4;
        double *p_walk = e->d_partials, *p_gen = p_walk + (size_t)e->rows_walk * NECI_ST_COUNT,
               *p_eval = p_gen + (size_t)e->rows_gen * NECI_ST_COUNT, *p_sing = p_eval + (size_t)e->rows_eval * NECI_ST_COUNT,
               *p_heavy = p_sing + (size_t)e->rows_sing * NECI_ST_COUNT;
        NG_DISPATCH(e, (k_walk<NW, SYS><<<e->rows_walk, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, e->K, A, p_walk)));
        NG_DISPATCH(e, (k_generate<NW, SYS><<<e->rows_gen, K1_GEN_BLOCK, gen_smem_bytes<NW>(), e->stream>>>(e->P, e->L, e->K, A, p_gen)));
        NG_DISPATCH(e, (k_generate_heavy<NW, SYS><<<e->rows_heavy, K1_GEN_BLOCK, (K1_GEN_BLOCK / 32) * sizeof(GenStage<NW>), e->stream>>>(e->P, e->L, e->SB, e->K, A, p_heavy)));
        NG_DISPATCH(e, (k_evaluate<NW, SYS><<<e->rows_eval, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, e->K, A, p_eval)));
        if (pchb) NG_DISPATCH(e, (k_singles<NW, SYS><<<e->rows_sing, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, e->K, A, p_sing)));
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev[2], e->stream));
    if (e->cfg.nranks > 1) {
        long long nrecv = 0;
        if (exchange(e, &nrecv, true)) return 1;
        A.n_recv = nrecv;
    }
    CK(cudaEventRecord(e->ev[3], e->stream));
    if (annihilation_phase(e, A, e->rows_spawn + e->rows_heavy)) return 1;
    return finish_iteration(e, stats_out);
}

int neci_gpu_iterate_host(neci_gpu_engine *e, int64_t *current_dets, int64_t *n, double *gd, double *go,
                          double tau, double diag_sft, int64_t iter, double *stats_out) {
    CK(cudaSetDevice(e->cfg.device));
    if (neci_gpu_upload_walkers(e, current_dets, *n, gd, go)) return 1;
    // Host mirror: when the host arrays are page-locked and mapped (neci_gpu_alloc_host), the kernels write every
    // change of the iteration through to them, and nothing is downloaded in bulk.  Not with a core or trial space
    // (their set-up kernels touch flags outside an iteration) and not for pageable memory: then the whole list
    // is copied back as before.
    long long *m_rec = nullptr; double *m_gd = nullptr, *m_go = nullptr;
    bool mirror = !e->cfg.t_semi_stochastic && !e->P.trial_ht && e->cfg.nranks >= 1;
    if (mirror && cudaHostGetDevicePointer((void **)&m_rec, (void *)current_dets, 0) != cudaSuccess) { mirror = false; cudaGetLastError(); }
    if (mirror && gd && cudaHostGetDevicePointer((void **)&m_gd, (void *)gd, 0) != cudaSuccess) { mirror = false; cudaGetLastError(); }
    if (mirror && go && cudaHostGetDevicePointer((void **)&m_go, (void *)go, 0) != cudaSuccess) { mirror = false; cudaGetLastError(); }
    if (!mirror) {
        if (neci_gpu_iterate(e, tau, diag_sft, iter, stats_out)) return 1;
        return neci_gpu_download_walkers(e, current_dets, n, gd, go);
    }
    e->L.h_rec = m_rec; e->L.h_gd = m_gd; e->L.h_go = m_go; e->L.h_W = e->W;
    const int rc = neci_gpu_iterate(e, tau, diag_sft, iter, stats_out);       // ends with a stream synchronisation
    e->L.h_rec = nullptr; e->L.h_gd = nullptr; e->L.h_go = nullptr;
    if (rc) return 1;
    *n = e->n_resident;
    return 0;
}

int neci_gpu_annihilate(neci_gpu_engine *e, const int64_t *spawned_parts, int64_t n_spawned, int64_t iter, double *stats_out) {
    if (begin_iteration(e)) return 1;
    if (n_spawned > e->cfg.max_spawned) return e->fail("n_spawned exceeds max_spawned");
    CK(cudaMemsetAsync(e->d_partials, 0, (size_t)(e->rows_spawn + e->rows_heavy) * NECI_ST_COUNT * 8, e->stream));
    CK(cudaMemsetAsync(e->d_partials + (size_t)(e->rows_total - e->rows_trial - e->rows_tau) * NECI_ST_COUNT, 0,
                       (size_t)(e->rows_trial + e->rows_tau) * NECI_ST_COUNT * 8, e->stream));
    CK(cudaMemcpyAsync(e->SB.recv, spawned_parts, (size_t)n_spawned * e->W * 8, cudaMemcpyHostToDevice, e->stream));
    IterArgs A; A.tau = 0; A.diag_sft = 0; A.iter = iter; A.n_recv = n_spawned; A.stamp = 0;
    for (int k = 0; k < 4; ++k) CK(cudaEventRecord(e->ev[k], e->stream));
    const bool semi = e->cfg.t_semi_stochastic;
    e->cfg.t_semi_stochastic = 0;                 // no determ_projection output to apply here
    const int rc = annihilation_phase(e, A, e->rows_spawn + e->rows_heavy);
    e->cfg.t_semi_stochastic = semi;
    if (rc) return 1;
    return finish_iteration(e, stats_out);
}

// -------------------------------------------------------------------------------
int neci_gpu_nccl_unique_id(uint8_t id_out[128]) {
    std::string err;
    if (!g_nccl.load(err)) { fprintf(stderr, "neci_gpu: %s\n", err.c_str()); return 1; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id_out, &id, 128);
    return 0;
}
int neci_gpu_nccl_init(neci_gpu_engine *e, const uint8_t id_in[128]) {
    if (!g_nccl.load(e->err)) return 1;
    CK(cudaSetDevice(e->cfg.device));
    ncclUniqueId id; memcpy(&id, id_in, 128);
    NCK(g_nccl.CommInitRank(&e->comm, e->cfg.nranks, id, e->cfg.rank));
    return 0;
}

// ---- peer-memory exchange wiring -----------------------------------------------------
int neci_gpu_p2p_handle(neci_gpu_engine *e, uint8_t handle_out[64]) {
    CK(cudaSetDevice(e->cfg.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    const int nr = e->cfg.nranks;
    if (!e->p2p_block) {
        e->p2p_seg_words = (size_t)2 * nr * e->SB.seg_cap * e->W;
        const size_t bytes = e->p2p_seg_words * 8 + (size_t)2 * nr * 8;
        CK(cudaMalloc(&e->p2p_block, bytes));
        CK(cudaMemset(e->p2p_block, 0, bytes));
        CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->p2p_block));
    memcpy(handle_out, &h, 64);
    return 0;
}
int neci_gpu_p2p_open(neci_gpu_engine *e, const uint8_t *handles) {
    CK(cudaSetDevice(e->cfg.device));
    const int nr = e->cfg.nranks;
    if (!e->p2p_block) return e->fail("neci_gpu_p2p_handle must be called first");
    e->p2p_peer_base.assign(nr, nullptr);
    std::vector<long long *> seg(nr); std::vector<unsigned long long *> mail(nr);
    for (int r = 0; r < nr; ++r) {
        void *base = e->p2p_block;
        if (r != e->cfg.rank) {
            cudaIpcMemHandle_t h; memcpy(&h, handles + (size_t)r * 64, 64);
            cudaError_t rc = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (rc != cudaSuccess) return e->fail("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(rc));
        }
        e->p2p_peer_base[r] = base;
        seg[r] = (long long *)base; mail[r] = (unsigned long long *)((long long *)base + e->p2p_seg_words);
    }
    e->X.peer_seg = e->upload(seg.data(), seg.size());
    e->X.peer_mail = e->upload(mail.data(), mail.size());
    e->X.my_seg = seg[e->cfg.rank]; e->X.my_mail = mail[e->cfg.rank];
    e->X.cnt_in = e->alloc<unsigned long long>((size_t)2 * nr);
    e->X.ticket = e->alloc<unsigned int>(1);
    e->SB.n_recv_dev = e->alloc<unsigned long long>(1);
    e->SB.push_cnt = e->alloc<unsigned long long>((size_t)nr * NG_PUSH_CNT_STRIDE);
    if (!e->X.peer_seg || !e->X.peer_mail || !e->X.cnt_in || !e->X.ticket || !e->SB.n_recv_dev || !e->SB.push_cnt) return e->fail("allocation failed");
    CK(cudaMemset(e->SB.push_cnt, 0, (size_t)nr * NG_PUSH_CNT_STRIDE * 8));
    e->SB.push_seg = e->X.peer_seg;
    CK(cudaMemset(e->X.ticket, 0, 4));
    CK(cudaMemset(e->SB.n_recv_dev, 0, 8));
    CK(cudaDeviceSynchronize());
    e->xseq = 0;
    e->p2p = true;
    if (const char *t = getenv("NECI_GPU_TIMING")) if (t[0] == '1') {
        e->x_timing = true;
        for (auto &v : e->x_ev) CK(cudaEventCreate(&v));
    }
    return 0;
}

int neci_gpu_block_populations(neci_gpu_engine *e, double *block_parts) {
    CK(cudaSetDevice(e->cfg.device));
    const int nb = e->cfg.balance_blocks;
    double *d = dalloc<double>(nb);
    if (!d) return e->fail("allocation failed");
    cudaMemsetAsync(d, 0, nb * 8, e->stream);
    if (e->nw == 1) k_block_pops<1><<<e->grid_generic, 256, 0, e->stream>>>(e->P, e->L, d);
    else k_block_pops<2><<<e->grid_generic, 256, 0, e->stream>>>(e->P, e->L, d);
    cudaMemcpyAsync(block_parts, d, nb * 8, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t rc = cudaStreamSynchronize(e->stream);
    cudaFree(d);
    if (rc != cudaSuccess) return e->fail("block population kernel failed: %s", cudaGetErrorString(rc));
    return 0;
}

int neci_gpu_rebalance(neci_gpu_engine *e, const int32_t *new_mapping) {
    // adjust_load_balance / move_block (load_balancer.fpp:178-512): install the new LoadBalanceMapping, ship
    // every determinant whose block changed owner to its new rank (same grouped send/recv as the spawn
    // exchange) and insert it there (AddNewHashDet recomputes H_ii and H_0i, as the receiver does in move_block).
    CK(cudaSetDevice(e->cfg.device));
    if (e->cfg.t_semi_stochastic && e->n_core_total > 0)
        return e->fail("should not be dynamically load-balancing with a fixed deterministic space (load_balancer.fpp:198-201)");
    for (int b = 0; b < e->cfg.balance_blocks; ++b)
        if (new_mapping[b] < 0 || new_mapping[b] >= e->cfg.nranks) return e->fail("new_mapping[%d] = %d out of range", b, new_mapping[b]);
    CK(cudaMemcpyAsync((void *)e->P.lb_mapping, new_mapping, (size_t)e->cfg.balance_blocks * 4, cudaMemcpyHostToDevice, e->stream));
    if (e->cfg.nranks == 1) { CK(cudaStreamSynchronize(e->stream)); return 0; }
    if (begin_iteration(e)) return 1;
    const int g = e->grid_generic;
    e->n_launch += 5;
    if (e->nw == 1) k_rebalance_pack<1><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB);
    else k_rebalance_pack<2><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB);
    CK(cudaGetLastError());
    long long nrecv = 0;
    if (exchange(e, &nrecv, false)) return 1;
    k_merge_free<<<64, 256, 0, e->stream>>>(e->L);
    k_merge_free_finish<<<1, 1, 0, e->stream>>>(e->L);
    k_iota_insert<<<g, 256, 0, e->stream>>>(e->L, e->SB, nrecv);
    double *p_ins = e->d_partials + (size_t)(e->rows_spawn + e->rows_heavy + e->rows_compress + e->rows_annih) * NECI_ST_COUNT;
    NG_DISPATCH(e, (k_insert<NW, SYS><<<g, NG_BLOCK, 0, e->stream>>>(e->P, e->L, e->SB, p_ins)));
    k_fix_counters<<<1, 1, 0, e->stream>>>(e->L);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_ctr, e->L.ctr, C_COUNT * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const long long errf = e->h_ctr[C_ERR];
    if (errf & 1) return e->fail("rebalance: moved determinants exceed the spawn buffer segment (move fewer blocks per call or raise max_spawned)");
    if (errf & 2) return e->fail("rebalance: main walker list overflow on the receiving rank");
    return 0;
}

// ---- measurement helpers ------------------------------------------------------------
int neci_gpu_timer_start(neci_gpu_engine *e) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventRecord(e->ev_t0, e->stream));
    return 0;
}
int neci_gpu_timer_stop(neci_gpu_engine *e, double *ms_out) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaEventRecord(e->ev_t1, e->stream));
    CK(cudaEventSynchronize(e->ev_t1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
    if (ms_out) *ms_out = (double)ms;
    return 0;
}
int64_t neci_gpu_launch_count(const neci_gpu_engine *e) { return e ? e->n_launch : 0; }
int neci_gpu_alloc_host(int64_t bytes, void **out) {
    if (!out || bytes < 0) return 1;
    return cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 8)) == cudaSuccess ? 0 : 1;
}
int neci_gpu_free_host(void *p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? 0 : 1; }

// ---- probes -----------------------------------------------------------------------
int neci_gpu_probe_det_node(neci_gpu_engine *e, int64_t n, const int64_t *iluts, int32_t *block_out, int32_t *node_out) {
    CK(cudaSetDevice(e->cfg.device));
    Scratch sc;
    long long *d_il = sc.get<long long>((size_t)n * e->nw); int *d_b = sc.get<int>(n), *d_n = sc.get<int>(n);
    if (!sc.ok) return e->fail("probe_det_node: device allocation failed (n = %lld)", (long long)n);
    CK(cudaMemcpy(d_il, iluts, (size_t)n * e->nw * 8, cudaMemcpyHostToDevice));
    const int grid = (int)std::max<long long>(1, std::min<long long>(e->grid_generic, (n + 255) / 256));
    if (e->nw == 1) k_probe_det_node<1><<<grid, 256, 0, e->stream>>>(e->P, d_il, n, d_b, d_n);
    else k_probe_det_node<2><<<grid, 256, 0, e->stream>>>(e->P, d_il, n, d_b, d_n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(block_out, d_b, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(node_out, d_n, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}
int neci_gpu_probe_helement(neci_gpu_engine *e, int64_t n, const int64_t *ii, const int64_t *ij, double *out) {
    CK(cudaSetDevice(e->cfg.device));
    Scratch sc;
    long long *d_i = sc.get<long long>((size_t)n * e->nw), *d_j = sc.get<long long>((size_t)n * e->nw); double *d_o = sc.get<double>(n);
    if (!sc.ok) return e->fail("probe_helement: device allocation failed (n = %lld)", (long long)n);
    CK(cudaMemcpy(d_i, ii, (size_t)n * e->nw * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_j, ij, (size_t)n * e->nw * 8, cudaMemcpyHostToDevice));
    const int grid = (int)std::max<long long>(1, std::min<long long>(e->grid_generic, (n + 255) / 256));
    NG_DISPATCH(e, (k_probe_helement<NW, SYS><<<grid, 256, 0, e->stream>>>(e->P, d_i, d_j, n, d_o)));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, d_o, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int neci_gpu_probe_gen_excit(neci_gpu_engine *e, int64_t n, const int64_t *iluts, const int32_t *attempt, int64_t iter,
                             int64_t *ilut_j_out, int32_t *ic_out, int32_t *ex_out, int32_t *parity_out,
                             double *pgen_out, double *hel_out) {
    CK(cudaSetDevice(e->cfg.device));
    Scratch sc;
    long long *d_il = sc.get<long long>((size_t)n * e->nw), *d_j = sc.get<long long>((size_t)n * e->nw);
    int *d_at = sc.get<int>(n), *d_ic = sc.get<int>(n), *d_ex = sc.get<int>(4 * n), *d_par = sc.get<int>(n);
    double *d_pg = sc.get<double>(n), *d_h = sc.get<double>(n);
    if (!sc.ok) return e->fail("probe_gen_excit: device allocation failed (n = %lld)", (long long)n);
    CK(cudaMemcpy(d_il, iluts, (size_t)n * e->nw * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_at, attempt, n * 4, cudaMemcpyHostToDevice));
    const int grid = (int)std::max<long long>(1, std::min<long long>(e->grid_generic, (n + 255) / 256));
    NG_DISPATCH(e, (k_probe_gen_excit<NW, SYS><<<grid, 256, 0, e->stream>>>(e->P, d_il, d_at, iter, n, d_j, d_ic, d_ex, d_par, d_pg, d_h)));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(ilut_j_out, d_j, (size_t)n * e->nw * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ic_out, d_ic, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ex_out, d_ex, 4 * n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(parity_out, d_par, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pgen_out, d_pg, n * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hel_out, d_h, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
