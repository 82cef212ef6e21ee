// xtb_api.cu -- the C-ABI of libxtb200.so (include/xtb200.h) and the small
// kernels around the tracking kernel (rng seeding, statistics, compaction,
// DFMA peak measurement).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/xtb200.h"
#include "xtb_math.cuh"
#include "xtb_libm.cuh"
#include "xtb_ops.h"
#include "xtb_state.cuh"
#include "xtb_rng.cuh"

extern "C" cudaError_t xtb_launch_track_fast(unsigned, const XtbTrackArgs*, int, int*, cudaStream_t);
extern "C" cudaError_t xtb_launch_track_exact(unsigned, const XtbTrackArgs*, int, int*, cudaStream_t);


static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CUDA_TRY(expr)                                                          \
    do {                                                                        \
        cudaError_t e_ = (expr);                                                \
        if (e_ != cudaSuccess) {                                                \
            snprintf(g_err, sizeof(g_err), "%s: %s", #expr, cudaGetErrorString(e_)); \
            return XTB_E_CUDA;                                                  \
        }                                                                       \
    } while (0)

// One lowered program resident on the device (see xtb_ops.h: FUSED and PLAIN).
struct xtb_program {
    size_t n_words = 0, n_tiles = 0;
    bool has_heavy = false;
    bool has_beam_mon = false;     // XTB_OP_BEAM_MON / XTB_OP_BEAM_PROFILE present
    bool has_quantum = false;      // a magnet body with radiation_flag 2 / 3 (random emission)
    bool has_qk = false;           // radiation_flag 3: needs the inverse-CDF tables
    bool has_fast_aperture = false;    // XTB_OP_RECT / XTB_OP_ELLIPSE (with or without drift prefix)
    uint64_t* d_prog = nullptr;          // IMAGE: tile k = its ops + one XTB_OP_END op (2 words)
    uint32_t* d_tile_off = nullptr;      // image offsets [n_tiles + 1]
    std::vector<uint32_t> elem_offset;   // host copy [n_elements + 1], XTB_NOT_ADDRESSABLE allowed
    std::vector<uint32_t> tile_off;      // host copy [n_tiles + 1], offsets in the op stream
    // tile holding word w of the op stream
    size_t tile_of(uint32_t w) const {
        size_t k = (size_t) (std::upper_bound(tile_off.begin(), tile_off.end(), w) - tile_off.begin());
        k = k ? k - 1 : 0;
        return std::min(k, n_tiles ? n_tiles - 1 : 0);
    }
};

struct xtb_lattice {
    int device;
    int n_sm;
    size_t n_elements;
    double line_length;
    xtb_program fused, plain;
    xtb_monitor_t* d_mon;
    xtb_last_turns_monitor_t* d_ltm;
    double* d_synrad_tables;      // quantum-kick model (xtb_lattice_set_synrad_tables)
};

extern "C" const char* xtb_last_error_string(void) { return g_err; }
extern "C" const char* xtb_version(void) { return "xtb200 0.3 (sm_100a)"; }
extern "C" int xtb_ops_abi_version(void) { return XTB_OPS_ABI_VERSION; }
extern "C" int64_t xtb_launch_count(void) { return g_launches.load(); }

// Validates the op stream and cuts it into tiles at addressable element boundaries.
static int program_prepare(xtb_program& G, const uint64_t* words, size_t n_words,
                           const uint32_t* elem_offset, size_t n_elements) {
    if (elem_offset[0] != 0 || elem_offset[n_elements] != n_words)
        return fail(XTB_E_INVALID, "elem_offset does not span the program");
    G.n_words = n_words;
    G.elem_offset.assign(elem_offset, elem_offset + n_elements + 1);
    G.tile_off.push_back(0);
    uint32_t w0 = 0;
    for (size_t e = 1; e <= n_elements; ++e) {
        const uint32_t w1 = elem_offset[e];
        if (w1 == XTB_NOT_ADDRESSABLE) continue;     // absorbed in the op that starts at w0
        if (w1 < w0 || w1 > n_words || (w0 & 1u)) return fail(XTB_E_INVALID, "bad element offsets");
        if (w1 - w0 > XTB_TILE_WORDS) return fail(XTB_E_INVALID, "element larger than a program tile");
        uint32_t pc = w0;
        while (pc < w1) {
            const uint64_t h = words[pc];
            const uint32_t nw = (uint32_t) ((h >> 16) & 0xffffu);
            const uint32_t op = (uint32_t) (h & 0xffu);
            if (nw < 2 || (nw & 1u) || pc + nw > w1) return fail(XTB_E_INVALID, "malformed op in program");
            if (op >= XTB_HEAVY_FIRST) G.has_heavy = true;
            if (op < XTB_GENERIC_FIRST && ((op & ~(uint32_t) XTB_OPBIT_DRIFT) == XTB_OP_RECT || (op & ~(uint32_t) XTB_OPBIT_DRIFT) == XTB_OP_ELLIPSE))
                G.has_fast_aperture = true;
            if (op == XTB_OP_BEAM_MON || op == XTB_OP_BEAM_PROFILE || op == XTB_OP_BEAM_STATS) G.has_beam_mon = true;
            if (op == XTB_OP_MAGNET_BODY && (((uint32_t) (h >> 32) >> 10) & 3u) >= 2u) G.has_quantum = true;
            if (op == XTB_OP_MAGNET_BODY && (((uint32_t) (h >> 32) >> 10) & 3u) == 3u) G.has_qk = true;
            pc += nw;
        }
        if (w1 - G.tile_off.back() > XTB_TILE_WORDS) G.tile_off.push_back(w0);
        w0 = w1;
    }
    G.tile_off.push_back((uint32_t) n_words);
    G.n_tiles = G.tile_off.size() - 1;
    return XTB_OK;
}

// Device image: the ops of each tile followed by the XTB_OP_END sentinel that ends the
// interpreter loop (xtb_interp.cuh), so tile k starts at image word tile_off[k] + 2 k.
static cudaError_t program_upload(xtb_program& G, const uint64_t* words) {
    std::vector<uint64_t> image;
    std::vector<uint32_t> image_off;
    image.reserve(G.n_words + 2 * G.n_tiles + 8);
    for (size_t k = 0; k < G.n_tiles; ++k) {
        image_off.push_back((uint32_t) image.size());
        image.insert(image.end(), words + G.tile_off[k], words + G.tile_off[k + 1]);
        image.push_back(XTB_HDR(XTB_OP_END, 0, 2, 0));
        image.push_back(0);
    }
    image_off.push_back((uint32_t) image.size());
    image.resize(image.size() + 8, 0);
    cudaError_t e = cudaMalloc(&G.d_prog, image.size() * sizeof(uint64_t));
    if (e == cudaSuccess)
        e = cudaMemcpy(G.d_prog, image.data(), image.size() * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&G.d_tile_off, image_off.size() * sizeof(uint32_t));
    if (e == cudaSuccess)
        e = cudaMemcpy(G.d_tile_off, image_off.data(), image_off.size() * sizeof(uint32_t),
                       cudaMemcpyHostToDevice);
    return e;
}

static void program_free(xtb_program& G) {
    if (G.d_prog) cudaFree(G.d_prog);
    if (G.d_tile_off) cudaFree(G.d_tile_off);
    G.d_prog = nullptr;
    G.d_tile_off = nullptr;
}

extern "C" int xtb_lattice_create(const uint64_t* fused_words, size_t n_fused_words,
                                  const uint32_t* fused_elem_offset,
                                  const uint64_t* plain_words, size_t n_plain_words,
                                  const uint32_t* plain_elem_offset, size_t n_elements,
                                  double line_length, int device, xtb_lattice_handle* out) {
    if (!out || (!plain_words && n_plain_words) || !plain_elem_offset)
        return fail(XTB_E_INVALID, "null argument");
    if (fused_words && !fused_elem_offset) return fail(XTB_E_INVALID, "null argument");
    xtb_lattice* L = new (std::nothrow) xtb_lattice();
    if (!L) return fail(XTB_E_NOMEM, "out of host memory");
    L->device = device;
    L->n_elements = n_elements;
    L->line_length = line_length;
    L->d_mon = nullptr;
    L->d_ltm = nullptr;
    L->d_synrad_tables = nullptr;
    int rc = program_prepare(L->plain, plain_words, n_plain_words, plain_elem_offset, n_elements);
    if (rc == XTB_OK) {
        for (size_t e = 0; e <= n_elements; ++e)
            if (plain_elem_offset[e] == XTB_NOT_ADDRESSABLE)
                rc = fail(XTB_E_INVALID, "every element of the plain program must be addressable");
    }
    if (rc == XTB_OK && fused_words)
        rc = program_prepare(L->fused, fused_words, n_fused_words, fused_elem_offset, n_elements);
    if (rc != XTB_OK) {
        delete L;
        return rc;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&L->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = program_upload(L->plain, plain_words);
    if (e == cudaSuccess && fused_words) e = program_upload(L->fused, fused_words);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "lattice upload: %s", cudaGetErrorString(e));
        program_free(L->plain);
        program_free(L->fused);
        delete L;
        return XTB_E_CUDA;
    }
    *out = L;
    return XTB_OK;
}

extern "C" int xtb_lattice_destroy(xtb_lattice_handle L) {
    if (!L) return XTB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(L->device);
    program_free(L->plain);
    program_free(L->fused);
    if (L->d_mon) cudaFree(L->d_mon);
    if (L->d_ltm) cudaFree(L->d_ltm);
    if (L->d_synrad_tables) cudaFree(L->d_synrad_tables);
    cudaSetDevice(prev);
    delete L;
    return XTB_OK;
}

extern "C" int xtb_lattice_set_inline_monitors(xtb_lattice_handle L, const xtb_monitor_t* mons,
                                               size_t n_mons,
                                               const xtb_last_turns_monitor_t* ltms, size_t n_ltms) {
    if (!L) return fail(XTB_E_INVALID, "null lattice");
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(L->device));
    if (L->d_mon) { cudaFree(L->d_mon);  L->d_mon = nullptr; }
    if (L->d_ltm) { cudaFree(L->d_ltm);  L->d_ltm = nullptr; }
    if (n_mons) {
        CUDA_TRY(cudaMalloc(&L->d_mon, n_mons * sizeof(xtb_monitor_t)));
        CUDA_TRY(cudaMemcpy(L->d_mon, mons, n_mons * sizeof(xtb_monitor_t), cudaMemcpyHostToDevice));
    }
    if (n_ltms) {
        CUDA_TRY(cudaMalloc(&L->d_ltm, n_ltms * sizeof(xtb_last_turns_monitor_t)));
        CUDA_TRY(cudaMemcpy(L->d_ltm, ltms, n_ltms * sizeof(xtb_last_turns_monitor_t),
                            cudaMemcpyHostToDevice));
    }
    cudaSetDevice(prev);
    return XTB_OK;
}

extern "C" int xtb_lattice_set_synrad_tables(xtb_lattice_handle L, const double* blob, size_t n_doubles) {
    if (!L || !blob || n_doubles < 8) return fail(XTB_E_INVALID, "null / short table blob");
    const size_t per_table = (size_t) blob[0] + (size_t) blob[1] + (size_t) blob[2];
    const size_t n_tables = (size_t) blob[4] + 3;      // N = 1 .. direct max, 64, 128, 256
    if (n_doubles != 8 + per_table * (1 + n_tables))
        return fail(XTB_E_INVALID, "table blob does not have the stated layout");
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(L->device));
    if (L->d_synrad_tables) { cudaFree(L->d_synrad_tables);  L->d_synrad_tables = nullptr; }
    CUDA_TRY(cudaMalloc(&L->d_synrad_tables, n_doubles * sizeof(double)));
    CUDA_TRY(cudaMemcpy(L->d_synrad_tables, blob, n_doubles * sizeof(double), cudaMemcpyHostToDevice));
    cudaSetDevice(prev);
    return XTB_OK;
}

extern "C" int xtb_track(xtb_lattice_handle L, const xtb_particles_t* particles,
                         int64_t num_turns, int32_t ele_start, int32_t num_ele_track,
                         int32_t flag_end_turn_actions, int32_t flag_reset_s_at_end_turn,
                         int32_t flag_monitor, const xtb_monitor_t* tbt_monitor,
                         uint64_t track_flags, double global_xy_limit,
                         uint32_t variant_flags, void* cuda_stream) {
    if (!L || !particles) return fail(XTB_E_INVALID, "null argument");
    if (track_flags & (1ull << XTB_FLAG_BACKTRACK))
        return fail(XTB_E_UNSUPPORTED, "XS_FLAG_BACKTRACK: backtracking is lowered on the host (create the lattice from "
                                       "the inverse maps in reverse order, lowering.lower_line(backtrack=True)); "
                                       "the kernel takes no such flag");
    if (track_flags & ((1ull << XTB_FLAG_SR_TAPER) | (1ull << XTB_FLAG_SR_KICK_SAME_AS_FIRST)))
        return fail(XTB_E_UNSUPPORTED, "single-particle twiss flags (SR_TAPER / SR_KICK_SAME_AS_FIRST)");
    if (ele_start < 0 || num_ele_track < 0
        || (size_t) ele_start + (size_t) num_ele_track > L->n_elements)
        return fail(XTB_E_INVALID, "element range outside the line");
    if (num_turns < 0 || num_turns > 0x7fffffff) return fail(XTB_E_INVALID, "bad num_turns");
    if (flag_monitor != 0 && !tbt_monitor) return fail(XTB_E_INVALID, "flag_monitor without monitor");
    if (particles->capacity <= 0) return fail(XTB_E_INVALID, "empty particles");
    if (particles->capacity >= (1ll << 31)) return fail(XTB_E_INVALID, "capacity >= 2^31 slots per launch");
    for (int f = 0; f < XTB_NUM_FIELDS; ++f)
        if (!particles->field[f]) return fail(XTB_E_INVALID, "null particle field pointer");
    if (num_turns == 0) return XTB_OK;

    // the fused program serves every launch whose element range falls on its op boundaries;
    // the element-by-element monitor needs the per-element hooks of the plain program
    const xtb_program* G = &L->plain;
    if (L->fused.d_prog && flag_monitor != 2 && !(variant_flags & XTB_VARIANT_PLAIN_PROGRAM)
        && L->fused.elem_offset[ele_start] != XTB_NOT_ADDRESSABLE
        && L->fused.elem_offset[ele_start + num_ele_track] != XTB_NOT_ADDRESSABLE)
        G = &L->fused;

    XtbTrackArgs a;
    memset(&a, 0, sizeof(a));
    a.prog = G->d_prog;
    a.tile_off = G->d_tile_off;
    a.inline_mon = L->d_mon;
    a.inline_ltm = L->d_ltm;
    a.part = *particles;
    if (tbt_monitor) a.mon = *tbt_monitor;
    {   // op-stream range -> tiles and image offsets (every tile before adds 2 sentinel words)
        const uint32_t w_start = G->elem_offset[ele_start];
        const uint32_t w_stop = G->elem_offset[ele_start + num_ele_track];
        const size_t t_first = G->tile_of(w_start);
        const size_t t_last = w_stop > w_start ? G->tile_of(w_stop - 1) : t_first;
        a.tile_first = (int32_t) t_first;
        a.tile_last = (int32_t) t_last;
        a.pc_start = w_start + 2u * (uint32_t) t_first;
        a.pc_stop = (w_stop > w_start ? w_stop : w_start) + 2u * (uint32_t) t_last;
    }
    a.num_turns = (int32_t) num_turns;
    a.num_ele_track = (uint32_t) num_ele_track;
    a.flag_end_turn_actions = flag_end_turn_actions;
    a.flag_reset_s = flag_reset_s_at_end_turn;
    a.flag_monitor = flag_monitor;
    a.ignore_global = (int32_t) ((track_flags >> XTB_FLAG_IGNORE_GLOBAL_APERTURE) & 1);
    a.ignore_local = (int32_t) ((track_flags >> XTB_FLAG_IGNORE_LOCAL_APERTURE) & 1);
    a.kill_cavity_kick = (int32_t) ((track_flags >> XTB_FLAG_KILL_CAVITY_KICK) & 1);
    a.rng_philox = (variant_flags & XTB_VARIANT_PHILOX) ? 1 : 0;
    a.aperture_prefilter = G->has_fast_aperture ? 1 : 0;
    a.line_length = L->line_length;
    a.global_xy_limit = global_xy_limit;
    a.synrad_tables = L->d_synrad_tables;
    if (G->has_qk && (variant_flags & XTB_VARIANT_SYNRAD) && !L->d_synrad_tables)
        return fail(XTB_E_INVALID, "quantum-kick radiation needs xtb_lattice_set_synrad_tables first");

    unsigned variant = 0;
    if (G->has_heavy || (variant_flags & XTB_VARIANT_SYNRAD)) variant |= 1u;
    if (variant_flags & XTB_VARIANT_SYNRAD) variant |= 2u;
    if (variant_flags & XTB_VARIANT_FREEZE_LONG) variant |= 4u;
    if (G->has_beam_mon) variant |= 8u;
    if (G->has_quantum) variant |= 16u;

    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != L->device) CUDA_TRY(cudaSetDevice(L->device));
    int n_launched = 0;
    cudaError_t e = (variant_flags & XTB_VARIANT_EXACT)
                        ? xtb_launch_track_exact(variant, &a, L->n_sm, &n_launched, (cudaStream_t) cuda_stream)
                        : xtb_launch_track_fast(variant, &a, L->n_sm, &n_launched, (cudaStream_t) cuda_stream);
    if (prev != L->device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "track kernel launch: %s", cudaGetErrorString(e));
        return e == cudaErrorNotSupported ? XTB_E_UNSUPPORTED : XTB_E_CUDA;
    }
    g_launches.fetch_add(n_launched);
    return XTB_OK;
}

// ---- RNG seeding: rng_set of xtrack/particles/rng_src/base_rng.h:45-62 ----
#define XTB_TAUS(s, a, b, c, d) ((((s) & (c)) << (d)) ^ ((((s) << (a)) ^ (s)) >> (b)))
__device__ __forceinline__ uint32_t xtb_rng_u32(uint32_t& s1, uint32_t& s2, uint32_t& s3, uint32_t& s4) {
    s1 = XTB_TAUS(s1, 13, 19, 4294967294u, 12);
    s2 = XTB_TAUS(s2, 2, 25, 4294967288u, 4);
    s3 = XTB_TAUS(s3, 3, 11, 4294967280u, 17);
    s4 = 1664525u * s4 + 1013904223u;
    return s1 ^ s2 ^ s3 ^ s4;
}

__global__ void xtb_rng_init_kernel(uint32_t* r1, uint32_t* r2, uint32_t* r3, uint32_t* r4,
                                    const uint32_t* __restrict__ seeds, int64_t n) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = seeds[i];
    uint32_t s1 = 69069u * s;
    if (s1 < 2) s1 += 2u;
    uint32_t s2 = 69069u * s1;
    if (s2 < 8) s2 += 8u;
    uint32_t s3 = 69069u * s2;
    if (s3 < 16) s3 += 16u;
    uint32_t s4 = 69069u * s3;
    for (int k = 0; k < 6; ++k) xtb_rng_u32(s1, s2, s3, s4);
    r1[i] = s1;  r2[i] = s2;  r3[i] = s3;  r4[i] = s4;
}

extern "C" int xtb_rng_init(const xtb_particles_t* p, const uint32_t* seeds_dev, int64_t n,
                            int device, void* cuda_stream) {
    if (!p || !seeds_dev || n < 0 || n > p->capacity) return fail(XTB_E_INVALID, "bad rng_init arguments");
    if (n == 0) return XTB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device) CUDA_TRY(cudaSetDevice(device));
    xtb_rng_init_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, (cudaStream_t) cuda_stream>>>(
        (uint32_t*) p->field[F_RNG_S1], (uint32_t*) p->field[F_RNG_S2],
        (uint32_t*) p->field[F_RNG_S3], (uint32_t*) p->field[F_RNG_S4], seeds_dev, n);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "rng init launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return XTB_OK;
}

// ---- per-GPU partial statistics (K5) -------------------------------------
__global__ void xtb_stats_kernel(xtb_particles_t p, xtb_stats_t* out) {
    // 2 counters + 6 first moments + 21 second moments, warp-shuffle then atomics
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    double acc[27];
    long long n_alive = 0, n_lost = 0;
    for (int k = 0; k < 27; ++k) acc[k] = 0.;
    const int64_t* st = (const int64_t*) p.field[F_STATE];
    const int coord[6] = {F_X, F_PX, F_Y, F_PY, F_ZETA, F_DELTA};
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.capacity; i += stride) {
        const int64_t s = st[i];
        if (s > 0) {
            n_alive++;
            double v[6];
            for (int k = 0; k < 6; ++k) v[k] = ((const double*) p.field[coord[k]])[i];
            int m = 6;
            for (int r = 0; r < 6; ++r) {
                acc[r] += v[r];
                for (int c = r; c < 6; ++c) acc[m++] += v[r] * v[c];
            }
        } else if (s > -999999999) {
            n_lost++;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        for (int k = 0; k < 27; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], off);
        n_alive += __shfl_down_sync(0xffffffffu, n_alive, off);
        n_lost += __shfl_down_sync(0xffffffffu, n_lost, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*) &out->n_alive, (unsigned long long) n_alive);
        atomicAdd((unsigned long long*) &out->n_lost, (unsigned long long) n_lost);
        for (int k = 0; k < 6; ++k) atomicAdd(&out->sum[k], acc[k]);
        for (int k = 0; k < 21; ++k) atomicAdd(&out->sum2[k], acc[6 + k]);
    }
}

extern "C" int xtb_reduce_stats(const xtb_particles_t* p, xtb_stats_t* out_dev, int device,
                                void* cuda_stream) {
    if (!p || !out_dev) return fail(XTB_E_INVALID, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device) CUDA_TRY(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t) cuda_stream;
    CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(xtb_stats_t), s));
    unsigned grid = (unsigned) ((p->capacity + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    xtb_stats_kernel<<<grid, 256, 0, s>>>(*p, out_dev);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "stats launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return XTB_OK;
}

__global__ void xtb_loss_hist_kernel(const int64_t* __restrict__ state,
                                     const int64_t* __restrict__ at_element, int64_t capacity,
                                     int64_t n_elements, unsigned long long* hist) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= capacity) return;
    const int64_t s = state[i];
    if (s <= 0 && s > -999999999) {
        int64_t e = at_element[i];
        if (e < 0) e = 0;
        if (e > n_elements) e = n_elements;
        atomicAdd(&hist[e], 1ull);
    }
}

extern "C" int xtb_loss_histogram(const xtb_particles_t* p, int64_t* hist_dev, int64_t n_elements,
                                  int device, void* cuda_stream) {
    if (!p || !hist_dev || n_elements < 0) return fail(XTB_E_INVALID, "bad argument");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device) CUDA_TRY(cudaSetDevice(device));
    xtb_loss_hist_kernel<<<(unsigned) ((p->capacity + 255) / 256), 256, 0, (cudaStream_t) cuda_stream>>>(
        (const int64_t*) p->field[F_STATE], (const int64_t*) p->field[F_AT_ELEMENT], p->capacity,
        n_elements, (unsigned long long*) hist_dev);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "loss histogram launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return XTB_OK;
}

// ---- stream compaction (replaces the CPU reorganize()) --------------------
// Stable three-way partition [active | lost | unallocated] of the slots:
//   pass 1: per-block counts of the three classes;
//   pass 2: single-block exclusive scan of the counts;
//   pass 3: each slot computes its destination (block offset + rank in block),
//           writes perm[dst] = src;
//   pass 4: gather every field through perm via the scratch buffer.
#define XTB_CB 256
__device__ __forceinline__ int xtb_class_of(int64_t s) { return s > 0 ? 0 : (s > -999999999 ? 1 : 2); }

__global__ void xtb_compact_count(const int64_t* __restrict__ state, int64_t n, int64_t* counts) {
    __shared__ int c[3];
    if (threadIdx.x < 3) c[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t) blockIdx.x * XTB_CB + threadIdx.x;
    if (i < n) atomicAdd(&c[xtb_class_of(state[i])], 1);
    __syncthreads();
    if (threadIdx.x < 3) counts[(int64_t) threadIdx.x * gridDim.x + blockIdx.x] = c[threadIdx.x];
}

__global__ void xtb_compact_scan(int64_t* counts, int64_t n_blocks, int64_t* totals) {
    // serial scan by one thread: 3 * n_blocks entries (n_blocks = capacity/256), class-major
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int k = 0; k < 3; ++k) {
            int64_t before = run;
            for (int64_t b = 0; b < n_blocks; ++b) {
                const int64_t v = counts[k * n_blocks + b];
                counts[k * n_blocks + b] = run;
                run += v;
            }
            if (k < 2) totals[k] = run - before;
        }
    }
}

__global__ void xtb_compact_perm(const int64_t* __restrict__ state, int64_t n,
                                 const int64_t* __restrict__ offsets, int64_t* perm) {
    __shared__ int rank[3][XTB_CB / 32];
    const int64_t i = (int64_t) blockIdx.x * XTB_CB + threadIdx.x;
    const int cls = i < n ? xtb_class_of(state[i]) : 3;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned masks[3];
    for (int k = 0; k < 3; ++k) masks[k] = __ballot_sync(0xffffffffu, cls == k);
    if (lane == 0) for (int k = 0; k < 3; ++k) rank[k][warp] = __popc(masks[k]);
    __syncthreads();
    if (i < n) {
        int r = __popc(masks[cls] & ((1u << lane) - 1u));
        for (unsigned w = 0; w < warp; ++w) r += rank[cls][w];
        perm[offsets[(int64_t) cls * gridDim.x + blockIdx.x] + r] = i;
    }
}

template <typename T>
__global__ void xtb_gather(const T* __restrict__ src, T* __restrict__ dst,
                           const int64_t* __restrict__ perm, int64_t n) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

extern "C" size_t xtb_compact_scratch_bytes(int64_t capacity) {
    const int64_t n_blocks = (capacity + XTB_CB - 1) / XTB_CB;
    return (size_t) capacity * 8 + (size_t) n_blocks * 3 * 8 + 64;
}

extern "C" int xtb_compact(const xtb_particles_t* p, int64_t* perm_dev, int64_t* counts_dev,
                           void* scratch_dev, int device, void* cuda_stream) {
    if (!p || !perm_dev || !counts_dev || !scratch_dev) return fail(XTB_E_INVALID, "null argument");
    const int64_t n = p->capacity;
    const int64_t n_blocks = (n + XTB_CB - 1) / XTB_CB;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device) CUDA_TRY(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t) cuda_stream;
    char* scratch = (char*) scratch_dev;
    int64_t* blk = (int64_t*) (scratch + (size_t) n * 8);
    const int64_t* state = (const int64_t*) p->field[F_STATE];
    xtb_compact_count<<<(unsigned) n_blocks, XTB_CB, 0, s>>>(state, n, blk);
    xtb_compact_scan<<<1, 32, 0, s>>>(blk, n_blocks, counts_dev);
    xtb_compact_perm<<<(unsigned) n_blocks, XTB_CB, 0, s>>>(state, n, blk, perm_dev);
    const unsigned g = (unsigned) ((n + 255) / 256);
    for (int f = 0; f < XTB_NUM_FIELDS; ++f) {
        if (f < XTB_N_F64 + XTB_N_I64) {
            xtb_gather<uint64_t><<<g, 256, 0, s>>>((const uint64_t*) p->field[f], (uint64_t*) scratch, perm_dev, n);
            cudaMemcpyAsync(p->field[f], scratch, (size_t) n * 8, cudaMemcpyDeviceToDevice, s);
        } else {
            xtb_gather<uint32_t><<<g, 256, 0, s>>>((const uint32_t*) p->field[f], (uint32_t*) scratch, perm_dev, n);
            cudaMemcpyAsync(p->field[f], scratch, (size_t) n * 4, cudaMemcpyDeviceToDevice, s);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "compaction launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(3 + XTB_NUM_FIELDS);
    return XTB_OK;
}

// ---- DFMA peak: register-resident FMA chains -------------------------------
__global__ void __launch_bounds__(256) xtb_dfma_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c);  a1 = fma(a1, b, c);  a2 = fma(a2, b, c);  a3 = fma(a3, b, c);
            a4 = fma(a4, b, c);  a5 = fma(a5, b, c);  a6 = fma(a6, b, c);  a7 = fma(a7, b, c);
        }
    }
    out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int xtb_measure_dfma_peak(int device, double seconds, double* flops_out) {
    if (!flops_out) return fail(XTB_E_INVALID, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(device));
    int n_sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    const unsigned grid = (unsigned) n_sm * 8u;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, (size_t) grid * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 2000;
    double best = 0., elapsed_total = 0.;
    xtb_dfma_kernel<<<grid, 256>>>(d_out, iters, 1.0);   // warm-up
    cudaDeviceSynchronize();
    // sustained measurement: repeat until `seconds` of kernel time accumulated, report the
    // rate over the LAST half (clocks settled under load) and keep the best single launch
    double last_half_flop = 0., last_half_time = 0.;
    while (elapsed_total < seconds) {
        cudaEventRecord(e0);
        xtb_dfma_kernel<<<grid, 256>>>(d_out, iters, 1.0);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(d_out);
            return fail(XTB_E_CUDA, "dfma kernel: %s", cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = (double) grid * 256. * (double) iters * 64. * 2.;
        const double rate = flop / (ms * 1e-3);
        if (rate > best) best = rate;
        elapsed_total += ms * 1e-3;
        if (elapsed_total > 0.5 * seconds) { last_half_flop += flop;  last_half_time += ms * 1e-3; }
        if (ms < 20.f) iters *= 2;
    }
    g_launches.fetch_add(1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    cudaSetDevice(prev);
    flops_out[0] = last_half_time > 0 ? last_half_flop / last_half_time : best;
    flops_out[1] = best;
    return XTB_OK;
}

// ---- self-test of the guard-free FP64 sequences (xtb_math.cuh) ------------------------
__device__ __forceinline__ uint64_t xtb_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// a double with a random significand and a binary exponent drawn from [-erange, erange]
__device__ __forceinline__ double xtb_rand_double(uint64_t r, int erange, bool with_sign) {
    const uint64_t mant = r & 0x000fffffffffffffull;
    const int e = (erange > 0) ? (int) ((r >> 52) % (uint64_t) (2 * erange + 1)) - erange : 0;
    const uint64_t sign = with_sign ? (r >> 63) << 63 : 0ull;
    return __longlong_as_double((long long) (sign | ((uint64_t) (1023 + e) << 52) | mant));
}
__global__ void xtb_selftest_math_kernel(int64_t n_per_thread, uint64_t seed, int erange,
                                         unsigned long long* mismatches) {
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad_rcp = 0, bad_sqrt = 0, bad_div = 0;
    for (int64_t i = 0; i < n_per_thread; ++i) {
        const uint64_t r0 = xtb_mix64(seed + tid * 0x100000001B3ull + (uint64_t) i * 0x9E3779B9ull);
        const uint64_t r1 = xtb_mix64(r0);
        const double b = xtb_rand_double(r0, erange, true);
        const double a = xtb_rand_double(r1, erange, true);
        const double p = fabs(b);
        if (__double_as_longlong(xtb_rcp(b)) != __double_as_longlong(1.0 / b)) bad_rcp++;
        if (__double_as_longlong(xtb_sqrt(p)) != __double_as_longlong(sqrt(p))) bad_sqrt++;
        if (__double_as_longlong(xtb_div(a, b)) != __double_as_longlong(a / b)) bad_div++;
    }
    if (bad_rcp) atomicAdd(&mismatches[0], bad_rcp);
    if (bad_sqrt) atomicAdd(&mismatches[1], bad_sqrt);
    if (bad_div) atomicAdd(&mismatches[2], bad_div);
}

extern "C" int xtb_selftest_math(int device, int64_t n_samples, uint64_t seed, int exponent_range,
                                 uint64_t* mismatches_out /* [3]: rcp, sqrt, div */) {
    if (!mismatches_out || n_samples <= 0) return fail(XTB_E_INVALID, "bad argument");
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 3 * sizeof(unsigned long long));
    const unsigned grid = 148u * 8u, block = 256u;
    const int64_t per_thread = (n_samples + (int64_t) grid * block - 1) / ((int64_t) grid * block);
    xtb_selftest_math_kernel<<<grid, block>>>(per_thread, seed, exponent_range, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "math self-test: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    for (int i = 0; i < 3; ++i) mismatches_out[i] = h[i];
    return XTB_OK;
}

// ---- self-test of the glibc-compatible sin / cos (xtb_libm.cuh) -------------------------------
__global__ void xtb_eval_libm_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = xtb_sin_glibc(x[i]);
    out[n + i] = xtb_cos_glibc(x[i]);
    out[2 * n + i] = xtb_exp_glibc(x[i]);
    out[3 * n + i] = xtb_expm1_glibc(x[i]);
    out[4 * n + i] = xtb_sinh_glibc(x[i]);
    out[5 * n + i] = xtb_cosh_glibc(x[i]);
}

extern "C" int xtb_eval_libm(int device, const double* x_host, int64_t n, double* out_host) {
    if (!x_host || !out_host || n <= 0) return fail(XTB_E_INVALID, "bad argument");
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(device));
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t) n * 7 * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, x_host, (size_t) n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        xtb_eval_libm_kernel<<<(unsigned) ((n + 255) / 256), 256>>>(d, n, d + n);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d + n, (size_t) n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "libm self-test: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return XTB_OK;
}

// ---- self-test of the counter-based generator (xtb_thick.cuh::philox4x32_10) ------------------
__global__ void xtb_eval_philox_kernel(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, int64_t n,
                                       uint32_t* __restrict__ out) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b[4];
    philox4x32_10(k0, k1, c0 + (uint32_t) i, c1, b);
    for (int j = 0; j < 4; ++j) out[4 * i + j] = b[j];
}

extern "C" int xtb_eval_philox(int device, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, int64_t n,
                               uint32_t* out_host) {
    if (!out_host || n <= 0) return fail(XTB_E_INVALID, "bad argument");
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(device));
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t) n * 4 * sizeof(uint32_t)));
    xtb_eval_philox_kernel<<<(unsigned) ((n + 255) / 256), 256>>>(k0, k1, c0, c1, n, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d, (size_t) n * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(XTB_E_CUDA, "philox self-test: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return XTB_OK;
}
