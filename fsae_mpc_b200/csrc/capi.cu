// C-ABI of fsae_mpc_b200 (see include/fsae_mpc_b200.h).  Host-side plumbing only:
// context, parameter/track tables, staging buffers, kernel launches.  No CPU compute path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fsae_mpc_b200.h"
#include "launch.h"
#include "staged.cuh"
#include "probe.cuh"
#include "dense_qp.cuh"
#include "closed_loop.cuh"
#include "reference.cuh"

using namespace fsae;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Pinned (page-locked) host buffer, grow-only.
struct PinBuf {
    char* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc((void**)&p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Helper threads that move data between the caller's PAGEABLE buffers and the pinned staging ring
// (MATLAB's mxArrays and plain numpy arrays are pageable: a cudaMemcpyAsync on them is staged by the
// driver on the calling thread and serialises with everything else).  A job is one memcpy; a latch
// counts a group of jobs down.
struct Latch {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0;
    void reset(int k) { std::lock_guard<std::mutex> l(mu); n = k; }
    void count_down() { std::lock_guard<std::mutex> l(mu); if (--n <= 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return n <= 0; }); }
};
struct CopyJob {
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0;
    Latch* done = nullptr;
};
struct CopyPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<CopyJob> q;
    bool stop = false;
    int device = 0;
    void start(int n, int dev) {
        device = dev;
        for (int i = 0; i < n; ++i) th.emplace_back([this] { run(); });
    }
    void run() {
        cudaSetDevice(device);
        for (;;) {
            CopyJob j;
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                j = q.front();
                q.pop_front();
            }
            if (j.bytes) memcpy(j.dst, j.src, j.bytes);
            if (j.done) j.done->count_down();
        }
    }
    void post(const CopyJob& j) {
        { std::lock_guard<std::mutex> l(mu); q.push_back(j); }
        cv.notify_one();
    }
    void shutdown() {
        { std::lock_guard<std::mutex> l(mu); stop = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
    ~CopyPool() { shutdown(); }
};

constexpr int FSAE_RING = 6;        // slots of the pinned staging ring

struct fsae_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    std::string err;
    int64_t launches = 0;
    int kernel_version = 2;     // 1 = shared-memory operator (cross-check), 2 = register-tiled
    double *tap_H = nullptr, *tap_g = nullptr, *tap_M = nullptr;   // device buffers of the debug taps (tests)
    fsae_params h_params[FSAE_MAX_PARAM_SETS];
    fsae_params* d_params = nullptr;
    DevTrack h_tracks[FSAE_MAX_TRACKS];
    double* d_coef[FSAE_MAX_TRACKS];
    double track_len[FSAE_MAX_TRACKS];
    DevTrack* d_tracks = nullptr;
    unsigned long long* d_counters = nullptr;
    // staging for *_host calls
    DevBuf in[8], out[12];
    struct Slab { cudaStream_t st; DevBuf buf; };
    std::vector<Slab> slabs;   // per-problem L2 slabs of the long-horizon kernels, one pool PER STREAM
    // pageable callers: pinned staging ring + copy threads (created on first use)
    PinBuf ring_in[FSAE_RING], ring_out[FSAE_RING];
    cudaEvent_t ev_ring[FSAE_RING] = {};
    CopyPool* copy_pool = nullptr;
    int copy_threads = 0;
    int staging_mode = 0;      // 0 auto (stage when a caller buffer is pageable), 1 never, 2 always
    int last_host_path = 0;    // 0 direct copies, 1 pinned staging ring (fsae_debug_last_host_path)
};

#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
            return FSAE_ERR_CUDA;                                                      \
        }                                                                              \
    } while (0)

extern "C" const char* fsae_version(void) { return "fsae_mpc_b200 0.1 (sm_100a)"; }

extern "C" void fsae_default_params(int model, fsae_params* p) {
    memset(p, 0, sizeof(*p));
    p->lr = 0.6183; p->lf = 0.8672; p->mass = 280.0; p->inertia = 200.0; p->grav = 9.81;
    p->pac_B = 12.56; p->pac_C = 1.38; p->pac_D = 1.60; p->pac_E = -0.58;
    const double Q[7] = {5, 250, 2000, 0, 0, 0, 0};
    for (int i = 0; i < 7; ++i) { p->Q[i] = Q[i]; p->Q_terminal[i] = Q[i] * 10; }
    p->R[0] = 10; p->R[1] = 10;
    if (model == FSAE_MODEL_DYNAMIC) {
        p->R_soft[0] = 1e8; p->R_soft[1] = 1e6; p->R_soft[2] = 1e6; p->R_soft[3] = 1e4;
        p->lin_scheme = FSAE_LIN_RK4;
    } else {
        p->R_soft[0] = 1e8;
        p->lin_scheme = FSAE_LIN_RK2;
    }
    p->u_lb[0] = -10; p->u_lb[1] = -0.4; p->u_ub[0] = 10; p->u_ub[1] = 0.4;
    p->vel_lb = 0; p->vel_ub = INFINITY;
    p->delta_lb = -0.4; p->delta_ub = 0.4;
    p->n_lb = -0.75; p->n_ub = 0.75;
    p->soft_far = 1e10;
    p->ay_max = 5.0;
    p->slip_max = 0.1;
    p->ac_max = 9.163; p->al_max = 10.0;
    p->max_iter = -1;             /* qpOASES: maxIter = -1 -> 5*(nV+nC) (qpOASES_options.m:40-41) */
    p->feas_tol = 1e-9;
    p->flat_eps = 1e-8;
}

// Frees everything a context owns (also on the failure paths of fsae_create).
static void release_ctx(fsae_ctx* ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    if (ctx->copy_pool) { delete ctx->copy_pool; ctx->copy_pool = nullptr; }
    for (auto& b : ctx->in) b.release();
    for (auto& b : ctx->out) b.release();
    for (auto& sl : ctx->slabs) sl.buf.release();
    ctx->slabs.clear();
    for (auto& b : ctx->ring_in) b.release();
    for (auto& b : ctx->ring_out) b.release();
    for (auto& e : ctx->ev_ring) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (int i = 0; i < FSAE_MAX_TRACKS; ++i)
        if (ctx->d_coef[i]) { cudaFree(ctx->d_coef[i]); ctx->d_coef[i] = nullptr; }
    if (ctx->d_params) cudaFree(ctx->d_params);
    if (ctx->d_tracks) cudaFree(ctx->d_tracks);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int fsae_create(fsae_ctx** out, int device) {
    if (!out) return FSAE_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= device || device < 0) return FSAE_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FSAE_ERR_CUDA;
    if (prop.major != 10) return FSAE_ERR_CUDA;     // built for sm_100a only; no fallback
    fsae_ctx* ctx = new fsae_ctx();
    ctx->device = device;
    memset(ctx->h_tracks, 0, sizeof(ctx->h_tracks));
    memset(ctx->d_coef, 0, sizeof(ctx->d_coef));
    auto fail = [&](const char* what) {
        fprintf(stderr, "fsae_create: %s failed: %s\n", what, cudaGetErrorString(cudaGetLastError()));
        release_ctx(ctx);                            // streams, events and buffers made so far
        return FSAE_ERR_CUDA;
    };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) return fail("stream2");
    if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) return fail("event2");
    for (auto& e : ctx->ev_ring)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail("ring events");
    if (cudaMalloc(&ctx->d_params, sizeof(fsae_params) * FSAE_MAX_PARAM_SETS) != cudaSuccess) return fail("malloc params");
    if (cudaMalloc(&ctx->d_tracks, sizeof(DevTrack) * FSAE_MAX_TRACKS) != cudaSuccess) return fail("malloc tracks");
    if (cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess) return fail("malloc counters");
    cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long));
    cudaMemset(ctx->d_tracks, 0, sizeof(DevTrack) * FSAE_MAX_TRACKS);
    for (int i = 0; i < FSAE_MAX_PARAM_SETS; ++i) fsae_default_params(FSAE_MODEL_KINEMATIC, &ctx->h_params[i]);
    if (cudaMemcpy(ctx->d_params, ctx->h_params, sizeof(ctx->h_params), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail("memcpy params");
    {
        // copy threads of the pageable-host path: FSAE_COPY_THREADS, default min(8, cores / 2)
        const char* e = getenv("FSAE_COPY_THREADS");
        int t = e ? atoi(e) : 0;
        if (t <= 0) { t = (int)std::thread::hardware_concurrency() / 2; if (t > 8) t = 8; }
        ctx->copy_threads = t < 1 ? 1 : (t > 16 ? 16 : t);
    }
    *out = ctx;
    return FSAE_OK;
}

extern "C" int fsae_destroy(fsae_ctx* ctx) {
    if (!ctx) return FSAE_ERR_ARG;
    release_ctx(ctx);
    return FSAE_OK;
}

extern "C" const char* fsae_last_error(const fsae_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int64_t fsae_launch_count(const fsae_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* fsae_stream(const fsae_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" float fsae_last_kernel_ms(const fsae_ctx* ctx) {
    if (!ctx || !ctx->ev_valid) return 0.f;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return 0.f;
    return ms;
}

extern "C" int fsae_set_params(fsae_ctx* ctx, int id, const fsae_params* p) {
    if (!ctx || !p || id < 0 || id >= FSAE_MAX_PARAM_SETS) return FSAE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->h_params[id] = *p;
    CK(cudaMemcpyAsync(ctx->d_params + id, &ctx->h_params[id], sizeof(fsae_params), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

extern "C" int fsae_set_track(fsae_ctx* ctx, int track_id, const double* x_spline, const double* y_spline,
                              int n_seg, double dl) {
    if (!ctx || !x_spline || !y_spline || track_id < 0 || track_id >= FSAE_MAX_TRACKS || n_seg <= 0 || !(dl > 0))
        return FSAE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // MATLAB [n_seg x 4] column-major  ->  per-segment [x P0..P3 | y P0..P3]
    std::vector<double> h((size_t)n_seg * 8);
    for (int i = 0; i < n_seg; ++i)
        for (int k = 0; k < 4; ++k) {
            h[(size_t)i * 8 + k] = x_spline[(size_t)k * n_seg + i];
            h[(size_t)i * 8 + 4 + k] = y_spline[(size_t)k * n_seg + i];
        }
    if (ctx->d_coef[track_id]) { cudaFree(ctx->d_coef[track_id]); ctx->d_coef[track_id] = nullptr; }
    CK(cudaMalloc(&ctx->d_coef[track_id], h.size() * sizeof(double)));
    CK(cudaMemcpy(ctx->d_coef[track_id], h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    ctx->h_tracks[track_id].coef = ctx->d_coef[track_id];
    ctx->h_tracks[track_id].n_seg = n_seg;
    ctx->h_tracks[track_id].dl = dl;
    ctx->track_len[track_id] = dl * n_seg;       // arclength_reparam.m:29,64: L = M * dl
    CK(cudaMemcpy(ctx->d_tracks + track_id, &ctx->h_tracks[track_id], sizeof(DevTrack), cudaMemcpyHostToDevice));
    return FSAE_OK;
}

static int check_ids(fsae_ctx* ctx, int B, const int32_t* track_id, const int32_t* param_id) {
    if (track_id) {
        for (int i = 0; i < B; ++i)
            if (track_id[i] < 0 || track_id[i] >= FSAE_MAX_TRACKS || !ctx->h_tracks[track_id[i]].coef) {
                ctx->err = "track_id refers to a track that was never set";
                return FSAE_ERR_ARG;
            }
    } else if (!ctx->h_tracks[0].coef) {
        ctx->err = "track 0 not set (fsae_set_track)";
        return FSAE_ERR_ARG;
    }
    if (param_id)
        for (int i = 0; i < B; ++i)
            if (param_id[i] < 0 || param_id[i] >= FSAE_MAX_PARAM_SETS) {
                ctx->err = "param_id out of range";
                return FSAE_ERR_ARG;
            }
    return FSAE_OK;
}

extern "C" int fsae_interpolate_curvature_host(fsae_ctx* ctx, int track_id, const double* s, int64_t n,
                                               double* kappa_out) {
    if (!ctx || !s || !kappa_out || n < 0 || track_id < 0 || track_id >= FSAE_MAX_TRACKS) return FSAE_ERR_ARG;
    if (!ctx->h_tracks[track_id].coef) { ctx->err = "track not set"; return FSAE_ERR_ARG; }
    if (n == 0) return FSAE_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->in[0].reserve(n * sizeof(double)));
    CK(ctx->out[0].reserve(n * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->in[0].p, s, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    curvature_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->h_tracks[track_id], (const double*)ctx->in[0].p, (long long)n, (double*)ctx->out[0].p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(kappa_out, ctx->out[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

extern "C" int fsae_obtain_reference_host(fsae_ctx* ctx, const double* plan_x, const double* plan_t, int N_s, double ds,
                                          const double* s0, int B, double dt, int N_t, double* x_ref) {
    if (!ctx || !plan_x || !plan_t || !s0 || !x_ref || N_s < 1 || B < 0 || N_t < 1 || !(ds > 0.0) || !(dt > 0.0)) return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t bx = (size_t)8 * N_s * sizeof(double), bt = (size_t)N_s * sizeof(double), bs = (size_t)B * sizeof(double);
    const size_t bo = (size_t)7 * N_t * B * sizeof(double);
    CK(ctx->in[0].reserve(bx));
    CK(ctx->in[1].reserve(bt));
    CK(ctx->in[2].reserve(bs));
    CK(ctx->out[0].reserve(bo));
    CK(cudaMemcpyAsync(ctx->in[0].p, plan_x, bx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->in[1].p, plan_t, bt, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->in[2].p, s0, bs, cudaMemcpyHostToDevice, ctx->stream));
    fsae::obtain_reference_kernel<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(
        (const double*)ctx->in[0].p, (const double*)ctx->in[1].p, N_s, ds, (const double*)ctx->in[2].p, B, dt, N_t,
        (double*)ctx->out[0].p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x_ref, ctx->out[0].p, bo, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

static int model_dims(int model, int& NX, int& NU, int& NS) {
    if (model == FSAE_MODEL_KINEMATIC) { NX = 5; NU = 2; NS = 1; return FSAE_OK; }
    if (model == FSAE_MODEL_DYNAMIC) { NX = 7; NU = 2; NS = 4; return FSAE_OK; }
    return FSAE_ERR_ARG;
}

// upload optional id arrays; returns device pointers (or nullptr)
static int upload_ids(fsae_ctx* ctx, int B, const int32_t* track_id, const int32_t* param_id,
                      const int32_t** d_tid, const int32_t** d_pid) {
    *d_tid = nullptr;
    *d_pid = nullptr;
    if (track_id) {
        CK(ctx->in[6].reserve((size_t)B * 4));
        CK(cudaMemcpyAsync(ctx->in[6].p, track_id, (size_t)B * 4, cudaMemcpyHostToDevice, ctx->stream));
        *d_tid = (const int32_t*)ctx->in[6].p;
    }
    if (param_id) {
        CK(ctx->in[7].reserve((size_t)B * 4));
        CK(cudaMemcpyAsync(ctx->in[7].p, param_id, (size_t)B * 4, cudaMemcpyHostToDevice, ctx->stream));
        *d_pid = (const int32_t*)ctx->in[7].p;
    }
    return FSAE_OK;
}

extern "C" int fsae_linearise_host(fsae_ctx* ctx, int model, int B, int N, double dt,
                                   const int32_t* track_id, const int32_t* param_id,
                                   const double* x_lin, const double* u_lin,
                                   double* A, double* Bm, double* d) {
    int NX, NU, NS;
    if (!ctx || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || N <= 0 || !x_lin || !u_lin || !A || !Bm || !d)
        return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    int rc = check_ids(ctx, B, track_id, param_id);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t T = (size_t)B * N;
    CK(ctx->in[2].reserve(T * NX * 8));
    CK(ctx->in[3].reserve(T * NU * 8));
    CK(ctx->out[0].reserve(T * NX * NX * 8));
    CK(ctx->out[1].reserve(T * NX * NU * 8));
    CK(ctx->out[2].reserve(T * NX * 8));
    const int32_t *d_tid, *d_pid;
    rc = upload_ids(ctx, B, track_id, param_id, &d_tid, &d_pid);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->in[2].p, x_lin, T * NX * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->in[3].p, u_lin, T * NU * 8, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)((T + 127) / 128);
    if (model == FSAE_MODEL_KINEMATIC)
        linearise_kernel<KinModel><<<grid, 128, 0, ctx->stream>>>(B, N, dt, d_tid, d_pid, ctx->d_tracks, ctx->d_params,
            (const double*)ctx->in[2].p, (const double*)ctx->in[3].p, (double*)ctx->out[0].p, (double*)ctx->out[1].p, (double*)ctx->out[2].p);
    else
        linearise_kernel<DynModel><<<grid, 128, 0, ctx->stream>>>(B, N, dt, d_tid, d_pid, ctx->d_tracks, ctx->d_params,
            (const double*)ctx->in[2].p, (const double*)ctx->in[3].p, (double*)ctx->out[0].p, (double*)ctx->out[1].p, (double*)ctx->out[2].p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(A, ctx->out[0].p, T * NX * NX * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(Bm, ctx->out[1].p, T * NX * NU * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(d, ctx->out[2].p, T * NX * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

extern "C" int fsae_condense_host(fsae_ctx* ctx, int model, int B, int N, double dt,
                                  const int32_t* track_id, const int32_t* param_id,
                                  const double* x0, const double* x_ref,
                                  const double* x_lin, const double* u_lin,
                                  double* H, double* f, double* xA, double* lbA, double* ubA,
                                  double* lb, double* ub,
                                  double* A_bar, double* B_bar, double* d_bar, double* cost_const) {
    int NX, NU, NS;
    if (!ctx || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || N <= 0 || N > 80 || !x0 || !x_ref || !x_lin || !u_lin)
        return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    int rc = check_ids(ctx, B, track_id, param_id);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const int nU = NU * N, nV = nU + NS, nXN = NX * N;
    const int nC = (model == FSAE_MODEL_KINEMATIC) ? 6 * N : 20 * N;
    const size_t sz_in[4] = {(size_t)B * NX * 8, (size_t)B * nXN * 8, (size_t)B * nXN * 8, (size_t)B * nU * 8};
    const double* src[4] = {x0, x_ref, x_lin, u_lin};
    for (int i = 0; i < 4; ++i) {
        CK(ctx->in[i].reserve(sz_in[i]));
        CK(cudaMemcpyAsync(ctx->in[i].p, src[i], sz_in[i], cudaMemcpyHostToDevice, ctx->stream));
    }
    const int32_t *d_tid, *d_pid;
    rc = upload_ids(ctx, B, track_id, param_id, &d_tid, &d_pid);
    if (rc) return rc;
    const size_t sz_out[11] = {(size_t)B * nV * nV * 8, (size_t)B * nV * 8, (size_t)B * nC * nV * 8,
                               (size_t)B * nC * 8, (size_t)B * nC * 8, (size_t)B * nV * 8, (size_t)B * nV * 8,
                               (size_t)B * nXN * NX * 8, (size_t)B * nXN * nV * 8, (size_t)B * nXN * 8, (size_t)B * 8};
    double* dst[11] = {H, f, xA, lbA, ubA, lb, ub, A_bar, B_bar, d_bar, cost_const};
    for (int i = 0; i < 11; ++i) CK(ctx->out[i].reserve(sz_out[i]));
    CondenseArgs a;
    a.B = B; a.N = N; a.dt = dt; a.track_id = d_tid; a.param_id = d_pid;
    a.tracks = ctx->d_tracks; a.params = ctx->d_params;
    a.x0 = (const double*)ctx->in[0].p; a.x_ref = (const double*)ctx->in[1].p;
    a.x_lin = (const double*)ctx->in[2].p; a.u_lin = (const double*)ctx->in[3].p;
    a.H = H ? (double*)ctx->out[0].p : nullptr;
    a.f = f ? (double*)ctx->out[1].p : nullptr;
    a.xA = xA ? (double*)ctx->out[2].p : nullptr;
    a.lbA = (lbA && ubA) ? (double*)ctx->out[3].p : nullptr;
    a.ubA = (lbA && ubA) ? (double*)ctx->out[4].p : nullptr;
    a.lb = (lb && ub) ? (double*)ctx->out[5].p : nullptr;
    a.ub = (lb && ub) ? (double*)ctx->out[6].p : nullptr;
    a.A_bar = (double*)ctx->out[7].p; a.B_bar = (double*)ctx->out[8].p; a.d_bar = (double*)ctx->out[9].p;
    a.cconst = cost_const ? (double*)ctx->out[10].p : nullptr;
    const size_t ad_bytes = (size_t)N * NX * NX * sizeof(double);      // the kernel's dynamic shared memory: A_k
    // static + dynamic shared memory of the dynamic model at long horizons exceeds the 48 KB default: opt in
    CK(cudaFuncSetAttribute(condense_kernel<DynModel, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 7 * 7 * 8));
    if (model == FSAE_MODEL_KINEMATIC) condense_kernel<KinModel, 80><<<B, 256, ad_bytes, ctx->stream>>>(a);
    else condense_kernel<DynModel, 80><<<B, 256, ad_bytes, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    for (int i = 0; i < 11; ++i)
        if (dst[i]) CK(cudaMemcpyAsync(dst[i], ctx->out[i].p, sz_out[i], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

// ------------------------------------------------------------------ fused step
// Slab pool of the long-horizon kernels, one per stream that ever launched them: concurrent _dev calls on
// different streams never share per-CTA slabs.  Growing a pool (cudaFree + cudaMalloc) synchronises the
// device; it happens on the first call of a stream and when a larger batch slice arrives.
static DevBuf& slab_pool(fsae_ctx* ctx, cudaStream_t st) {
    for (auto& sl : ctx->slabs)
        if (sl.st == st) return sl.buf;
    ctx->slabs.push_back({st, DevBuf()});
    return ctx->slabs.back().buf;
}

// One launch, or -- for kernels that keep part of their state in a per-problem global slab (long horizons: the
// slab stays L2-resident with <= 148 CTAs in flight) -- slices of SLICE problems so that the slab pool is bounded.
typedef cudaError_t (*fused_launch_fn)(const BatchArgs&, cudaStream_t, int);
static int launch_fused(fsae_ctx* ctx, BatchArgs a, cudaStream_t st, fused_launch_fn fn, int variant, size_t slab_doubles,
                        int NX, int NU, int NS, int nC) {
    if (slab_doubles == 0) {
        CK(fn(a, st, variant));
        ctx->launches++;
        return FSAE_OK;
    }
    constexpr int SLICE = 2048;
    DevBuf& pool = slab_pool(ctx, st);
    const int nsl = a.B < SLICE ? a.B : SLICE;
    CK(pool.reserve((size_t)nsl * slab_doubles * sizeof(double)));
    const int B = a.B, N = a.N, nU = NU * N, nV = nU + NS;
    for (int lo = 0; lo < B; lo += SLICE) {
        BatchArgs c = a;
        c.B = (lo + SLICE <= B) ? SLICE : B - lo;
        c.m_scratch = (double*)pool.p;
        if (a.track_id) c.track_id = a.track_id + lo;
        if (a.param_id) c.param_id = a.param_id + lo;
        c.x0 = a.x0 + (size_t)lo * NX; c.x_ref = a.x_ref + (size_t)lo * NX * N;
        c.x_lin = a.x_lin + (size_t)lo * NX * N; c.u_lin = a.u_lin + (size_t)lo * NU * N;
        c.u_opt = a.u_opt + (size_t)lo * nU; c.x_opt = a.x_opt + (size_t)lo * NX * N;
        c.exitflag = a.exitflag + lo; c.fval = a.fval + lo; c.slack_opt = a.slack_opt + (size_t)lo * NS;
        if (a.iters) c.iters = a.iters + lo;
        if (a.wsB) c.wsB = a.wsB + (size_t)lo * nV;
        if (a.wsC) c.wsC = a.wsC + (size_t)lo * nC;
        CK(fn(c, st, variant));
        ctx->launches++;
    }
    return FSAE_OK;
}
#ifdef FSAE_XCHECK
static cudaError_t launch_v1_20(const BatchArgs& a, cudaStream_t st, int) { return launch_v1_kin(20, a, st); }
static cudaError_t launch_v1_40(const BatchArgs& a, cudaStream_t st, int) { return launch_v1_kin(40, a, st); }
static cudaError_t launch_v1_80(const BatchArgs& a, cudaStream_t st, int) { return launch_v1_kin(80, a, st); }
#endif

extern "C" int fsae_ltvmpc_dev(fsae_ctx* ctx, int model, int B, int N, double dt,
                               const int32_t* track_id, const int32_t* param_id,
                               const double* x0, const double* x_ref,
                               const double* x_lin, const double* u_lin,
                               double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                               double* slack_opt, int32_t* iters,
                               int8_t* workingSetB, int8_t* workingSetC, void* stream) {
    int NX, NU, NS;
    if (!ctx || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || !x0 || !x_ref || !x_lin || !u_lin ||
        !u_opt || !x_opt || !exitflag || !fval || !slack_opt)
        return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    // per-problem ids live on the device and cannot be checked here; the default track can
    if (!track_id && !ctx->h_tracks[0].coef) { ctx->err = "track 0 not set (fsae_set_track)"; return FSAE_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    BatchArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.N = N; a.dt = dt; a.track_id = track_id; a.param_id = param_id;
    a.x0 = x0; a.x_ref = x_ref; a.x_lin = x_lin; a.u_lin = u_lin;
    a.u_opt = u_opt; a.x_opt = x_opt; a.exitflag = exitflag; a.fval = fval; a.slack_opt = slack_opt;
    a.iters = iters; a.wsB = workingSetB; a.wsC = workingSetC;
    a.tracks = ctx->d_tracks; a.params = ctx->d_params;
    a.counters = ctx->d_counters;
    a.dbg_H = ctx->tap_H; a.dbg_g = ctx->tap_g; a.dbg_M = ctx->tap_M;
    CK(cudaEventRecord(ctx->ev0, st));
    int rc = FSAE_ERR_UNSUPPORTED;
    const int kv = ctx->kernel_version;     // 2 = product kernel; others only in the cross-check build
    const int nC = (model == FSAE_MODEL_KINEMATIC) ? 6 * N : 20 * N;
    // Runtime horizon: the reference takes any N_steps = length(x_ref) (ltvmpc_kinetmatic_curvilinear.m:17).  The
    // kernels are compiled for horizon capacities 20 / 40 / 80; a problem of N steps runs on the next capacity,
    // its remaining steps padded inside the kernel (zero state cost, rows disabled).
    if (N < 1) { ctx->err = "fused step: horizon must be >= 1"; return FSAE_ERR_ARG; }
    if (model == FSAE_MODEL_KINEMATIC) {
#ifdef FSAE_XCHECK
        if (kv == 1 && (N == 20 || N == 40 || N == 80))
            rc = launch_fused(ctx, a, st, N == 20 ? launch_v1_20 : (N == 40 ? launch_v1_40 : launch_v1_80), kv,
                              N == 80 ? slab_v1_kin80() : 0, NX, NU, NS, nC);
        else
#endif
        if (N <= 20) rc = launch_fused(ctx, a, st, launch_kin20, kv, 0, NX, NU, NS, nC);
        else if (N <= 40) rc = launch_fused(ctx, a, st, launch_kin40, kv, 0, NX, NU, NS, nC);
        else if (N <= 80) rc = launch_fused(ctx, a, st, launch_kin80, kv, slab_kin80(), NX, NU, NS, nC);
        else ctx->err = "kinematic fused step: horizon must be <= 80 (FSAE_MAX_HORIZON)";
    } else {
        if (N <= 20) rc = launch_fused(ctx, a, st, launch_dyn20, kv, 0, NX, NU, NS, nC);
        else if (N <= 40) rc = launch_fused(ctx, a, st, launch_dyn40, kv, 0, NX, NU, NS, nC);
        else if (N <= 80) rc = launch_fused(ctx, a, st, launch_dyn80, kv, slab_dyn80(), NX, NU, NS, nC);
        else ctx->err = "dynamic fused step: horizon must be <= 80 (FSAE_MAX_HORIZON)";
    }
    if (rc != FSAE_OK) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
    return FSAE_OK;
}

// true if a caller buffer is ordinary pageable memory (not cudaHostAlloc'ed / cudaHostRegister'ed)
static bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

extern "C" int fsae_ltvmpc_host(fsae_ctx* ctx, int model, int B, int N, double dt,
                                const int32_t* track_id, const int32_t* param_id,
                                const double* x0, const double* x_ref,
                                const double* x_lin, const double* u_lin,
                                double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                                double* slack_opt, int32_t* iters,
                                int8_t* workingSetB, int8_t* workingSetC) {
    int NX, NU, NS;
    if (!ctx || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || !x0 || !x_ref || !x_lin || !u_lin ||
        !u_opt || !x_opt || !exitflag || !fval || !slack_opt)
        return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    int rc = check_ids(ctx, B, track_id, param_id);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const int nU = NU * N, nV = nU + NS, nXN = NX * N;
    const int nC = (model == FSAE_MODEL_KINEMATIC) ? 6 * N : 20 * N;
    // Pipelined in chunks over two streams: the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap
    // the kernel of chunk c.  Pinned caller buffers are copied directly.  PAGEABLE caller buffers (MATLAB
    // mxArrays, plain numpy) go through a pinned staging ring filled / drained by helper threads, so that the
    // GPU copies stay asynchronous and the host memcpy's overlap the kernels too.
    const size_t per_in[4] = {(size_t)NX * 8, (size_t)nXN * 8, (size_t)nXN * 8, (size_t)nU * 8};
    const char* src[4] = {(const char*)x0, (const char*)x_ref, (const char*)x_lin, (const char*)u_lin};
    for (int i = 0; i < 4; ++i) CK(ctx->in[i].reserve(per_in[i] * B));
    const int32_t *d_tid, *d_pid;
    rc = upload_ids(ctx, B, track_id, param_id, &d_tid, &d_pid);
    if (rc) return rc;
    const size_t per_out[8] = {(size_t)nU * 8, (size_t)nXN * 8, 4, 8, (size_t)NS * 8, 4, (size_t)nV, (size_t)nC};
    char* dst[8] = {(char*)u_opt, (char*)x_opt, (char*)exitflag, (char*)fval, (char*)slack_opt, (char*)iters,
                    (char*)workingSetB, (char*)workingSetC};
    for (int i = 0; i < 8; ++i) CK(ctx->out[i].reserve(per_out[i] * B));
    bool staged = false;
    if (ctx->staging_mode != 1 && B >= 2048) {
        staged = ctx->staging_mode == 2;
        for (int i = 0; i < 4 && !staged; ++i) staged = is_pageable(src[i]);
        for (int i = 0; i < 8 && !staged; ++i) staged = is_pageable(dst[i]);
    }
    ctx->last_host_path = staged ? 1 : 0;
    // chunk size: pinned callers 1/8 of a large batch; staged callers 4096 problems (ring slot ~ 16 + 10 MB at N = 40)
    // staged callers: chunks of ~16 MB of input (4096 problems at N = 40, 8192 at N = 20), at least four per batch
    int per_staged = (int)(((size_t)16 << 20) / (per_in[0] + per_in[1] + per_in[2] + per_in[3]));
    per_staged = per_staged < 1024 ? 1024 : (per_staged / 1024) * 1024;
    if (per_staged > (B + 3) / 4) per_staged = (B + 3) / 4;
    const int per = staged ? per_staged : (B >= 16384 ? (B + 7) / 8 : B);
    // chunk boundaries.  Staged callers: the first and last chunks are short (1/4, 1/2 of a chunk), because the
    // copy-in of the first chunk and the copy-out of the last one cannot overlap any kernel.
    std::vector<int> cut{0};
    if (staged && B >= 4 * per) {
        const int head[2] = {per / 4, per / 2};
        for (int h : head) cut.push_back(cut.back() + h);
        const int tail_sz = per / 2 + per / 4;
        while (B - cut.back() - tail_sz > per) cut.push_back(cut.back() + per);
        const int rest = B - cut.back();                  // split the remainder: body | 1/2 | 1/4
        if (rest > tail_sz) cut.push_back(B - tail_sz);
        cut.push_back(B - per / 4);
        cut.push_back(B);
    } else {
        while (cut.back() < B) cut.push_back(cut.back() + per < B ? cut.back() + per : B);
    }
    const int nchunk = (int)cut.size() - 1;
    if (nchunk > 1) {
        // stream2 starts after whatever is already queued on the main stream (ids upload)
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    }
    size_t off_in[5] = {0, 0, 0, 0, 0}, off_out[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) off_in[i + 1] = off_in[i] + ((per_in[i] * per + 63) & ~(size_t)63);
    for (int i = 0; i < 8; ++i) off_out[i + 1] = off_out[i] + ((per_out[i] * per + 63) & ~(size_t)63);
    std::vector<Latch> lat_in, lat_out;
    if (staged) {
        if (!ctx->copy_pool) {
            ctx->copy_pool = new CopyPool();
            ctx->copy_pool->start(ctx->copy_threads, ctx->device);
        }
        for (int r = 0; r < FSAE_RING; ++r) {
            CK(ctx->ring_in[r].reserve(off_in[4]));
            CK(ctx->ring_out[r].reserve(off_out[8]));
        }
        lat_in = std::vector<Latch>(nchunk);
        lat_out = std::vector<Latch>(nchunk);
    }
    const int T = ctx->copy_threads;
    auto chunk_n = [&](int c) { return cut[c + 1] - cut[c]; };
    // an array is split over the copy threads in pieces of at least 256 KB (small arrays: one job)
    auto pieces = [&](size_t bytes) { const size_t k = (bytes + (256u << 10) - 1) / (256u << 10); return (int)(k < 1 ? 1 : (k > (size_t)T ? (size_t)T : k)); };
    // caller -> ring slot, split over the copy threads
    auto post_in = [&](int c) {
        const int lo = cut[c], n = chunk_n(c), slot = c % FSAE_RING;
        int njobs = 0;
        for (int i = 0; i < 4; ++i) njobs += pieces(per_in[i] * n);
        lat_in[c].reset(njobs);
        for (int i = 0; i < 4; ++i) {
            const size_t tot = per_in[i] * n;
            const int np = pieces(tot);
            const size_t piece = ((tot + np - 1) / np + 63) & ~(size_t)63;
            for (int t = 0; t < np; ++t) {
                const size_t o = (size_t)t * piece, len = o >= tot ? 0 : (o + piece <= tot ? piece : tot - o);
                CopyJob j;
                j.dst = ctx->ring_in[slot].p + off_in[i] + o;
                j.src = src[i] + per_in[i] * lo + o;
                j.bytes = len;
                j.done = &lat_in[c];
                ctx->copy_pool->post(j);
            }
        }
    };
    // ring slot -> caller; posted by the main thread once the chunk's D2H copies have completed, so every job in
    // the queue is runnable (a job that waited on an event would block the copy-in jobs queued behind it)
    auto post_out = [&](int c) {
        const int lo = cut[c], n = chunk_n(c), slot = c % FSAE_RING;
        int njobs = 0;
        for (int i = 0; i < 8; ++i) njobs += dst[i] ? pieces(per_out[i] * n) : 0;
        lat_out[c].reset(njobs);
        for (int i = 0; i < 8; ++i) {
            if (!dst[i]) continue;
            const size_t tot = per_out[i] * n;
            const int np = pieces(tot);
            const size_t piece = ((tot + np - 1) / np + 63) & ~(size_t)63;
            for (int t = 0; t < np; ++t) {
                const size_t o = (size_t)t * piece, len = o >= tot ? 0 : (o + piece <= tot ? piece : tot - o);
                CopyJob j;
                j.dst = dst[i] + per_out[i] * lo + o;
                j.src = ctx->ring_out[slot].p + off_out[i] + o;
                j.bytes = len;
                j.done = &lat_out[c];
                ctx->copy_pool->post(j);
            }
        }
    };
    constexpr int AHEAD = 3;            // chunks staged ahead of the one being enqueued
    static_assert(AHEAD < FSAE_RING, "ring depth");
    int fail_rc = FSAE_OK;
    if (staged)
        for (int c = 0; c < AHEAD && c < nchunk; ++c) post_in(c);
    for (int c = 0; c < nchunk; ++c) {
        const int lo = cut[c], n = chunk_n(c), slot = c % FSAE_RING;
        cudaStream_t st = (c & 1) ? ctx->stream2 : ctx->stream;
        if (staged) lat_in[c].wait();
        if (fail_rc == FSAE_OK) {
            auto ck = [&](cudaError_t e, const char* what) {
                if (e != cudaSuccess && fail_rc == FSAE_OK) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); fail_rc = FSAE_ERR_CUDA; }
            };
            for (int i = 0; i < 4; ++i)
                ck(cudaMemcpyAsync((char*)ctx->in[i].p + per_in[i] * lo,
                                   staged ? ctx->ring_in[slot].p + off_in[i] : src[i] + per_in[i] * lo, per_in[i] * n,
                                   cudaMemcpyHostToDevice, st), "H2D");
            auto o = [&](int i) { return (char*)ctx->out[i].p + per_out[i] * lo; };
            if (fail_rc == FSAE_OK) {
                rc = fsae_ltvmpc_dev(ctx, model, n, N, dt, d_tid ? d_tid + lo : nullptr, d_pid ? d_pid + lo : nullptr,
                                     (const double*)((char*)ctx->in[0].p + per_in[0] * lo),
                                     (const double*)((char*)ctx->in[1].p + per_in[1] * lo),
                                     (const double*)((char*)ctx->in[2].p + per_in[2] * lo),
                                     (const double*)((char*)ctx->in[3].p + per_in[3] * lo),
                                     (double*)o(0), (double*)o(1), (int32_t*)o(2), (double*)o(3), (double*)o(4),
                                     iters ? (int32_t*)o(5) : nullptr, workingSetB ? (int8_t*)o(6) : nullptr,
                                     workingSetC ? (int8_t*)o(7) : nullptr, st);
                if (rc) fail_rc = rc;
            }
            for (int i = 0; i < 8 && fail_rc == FSAE_OK; ++i)
                if (dst[i])
                    ck(cudaMemcpyAsync(staged ? ctx->ring_out[slot].p + off_out[i] : dst[i] + per_out[i] * lo, o(i),
                                       per_out[i] * n, cudaMemcpyDeviceToHost, st), "D2H");
        }
        if (staged) {
            cudaEventRecord(ctx->ev_ring[slot], st);
            // chunk c is queued on the GPU: now deliver chunk c - 1 (its event has fired or fires soon) ...
            if (c >= 1) {
                cudaEventSynchronize(ctx->ev_ring[(c - 1) % FSAE_RING]);
                post_out(c - 1);
            }
            // ... and stage chunk c + AHEAD, whose ring slot was last used by chunk c + AHEAD - RING
            const int nx = c + AHEAD;
            if (nx < nchunk) {
                if (nx - FSAE_RING >= 0) lat_out[nx - FSAE_RING].wait();
                post_in(nx);
            }
        }
    }
    if (staged) {
        cudaEventSynchronize(ctx->ev_ring[(nchunk - 1) % FSAE_RING]);
        post_out(nchunk - 1);
        for (int c = 0; c < nchunk; ++c) lat_out[c].wait();
    }
    if (nchunk > 1) {
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return fail_rc;
}

// Host-path selection of the _host entry points (tests / bench): 0 = automatic (pinned staging ring when a
// caller buffer is pageable), 1 = always direct copies, 2 = always the staging ring.  Returns the previous mode.
extern "C" int fsae_set_host_staging(fsae_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return FSAE_ERR_ARG;
    const int old = ctx->staging_mode;
    ctx->staging_mode = mode;
    return old;
}
// 1 if the most recent fsae_ltvmpc_host call went through the staging ring, 0 if it copied directly
extern "C" int fsae_last_host_path(const fsae_ctx* ctx) { return ctx ? ctx->last_host_path : -1; }

// ------------------------------------------------------------------ device pool
// One host thread drives every GPU of the box: the batch is split into contiguous shards (fsae_shard_range,
// the same rule as fsae_mpc_b200/sharding.py) and each shard runs through fsae_ltvmpc_host on its own context,
// on a helper thread per device for the duration of the call.  No collective, no peer traffic: the problems
// are independent.  This is what a single MATLAB process (one MEX handle) uses to reach all 8 GPUs.
struct fsae_pool {
    std::vector<fsae_ctx*> ctx;
    std::string err;
};

extern "C" void fsae_shard_range(int64_t total, int rank, int world, int64_t* lo, int64_t* hi) {
    const int64_t base = total / world, rem = total % world;
    const int64_t l = rank * base + (rank < rem ? rank : rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < rem ? 1 : 0);
}

extern "C" int fsae_pool_create(fsae_pool** out, const int* devices, int n_devices) {
    if (!out) return FSAE_ERR_ARG;
    *out = nullptr;
    std::vector<int> devs;
    if (devices) {
        if (n_devices < 1) return FSAE_ERR_ARG;
        devs.assign(devices, devices + n_devices);
    } else {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return FSAE_ERR_CUDA;
        if (n_devices > 0 && n_devices < n) n = n_devices;
        for (int i = 0; i < n; ++i) devs.push_back(i);
    }
    fsae_pool* p = new fsae_pool();
    for (int d : devs) {
        fsae_ctx* c = nullptr;
        const int rc = fsae_create(&c, d);
        if (rc != FSAE_OK) {
            for (auto* k : p->ctx) fsae_destroy(k);
            delete p;
            return rc;
        }
        p->ctx.push_back(c);
    }
    *out = p;
    return FSAE_OK;
}

extern "C" int fsae_pool_destroy(fsae_pool* p) {
    if (!p) return FSAE_ERR_ARG;
    for (auto* c : p->ctx) fsae_destroy(c);
    delete p;
    return FSAE_OK;
}
extern "C" int fsae_pool_size(const fsae_pool* p) { return p ? (int)p->ctx.size() : 0; }
extern "C" fsae_ctx* fsae_pool_ctx(fsae_pool* p, int i) { return (p && i >= 0 && i < (int)p->ctx.size()) ? p->ctx[i] : nullptr; }
extern "C" const char* fsae_pool_last_error(const fsae_pool* p) { return p ? p->err.c_str() : "null pool"; }

extern "C" int fsae_pool_set_track(fsae_pool* p, int track_id, const double* x_spline, const double* y_spline, int n_seg, double dl) {
    if (!p) return FSAE_ERR_ARG;
    for (auto* c : p->ctx) {
        const int rc = fsae_set_track(c, track_id, x_spline, y_spline, n_seg, dl);
        if (rc != FSAE_OK) { p->err = fsae_last_error(c); return rc; }
    }
    return FSAE_OK;
}
extern "C" int fsae_pool_set_params(fsae_pool* p, int id, const fsae_params* prm) {
    if (!p) return FSAE_ERR_ARG;
    for (auto* c : p->ctx) {
        const int rc = fsae_set_params(c, id, prm);
        if (rc != FSAE_OK) { p->err = fsae_last_error(c); return rc; }
    }
    return FSAE_OK;
}

extern "C" int fsae_ltvmpc_host_pool(fsae_pool* p, int model, int B, int N, double dt,
                                     const int32_t* track_id, const int32_t* param_id,
                                     const double* x0, const double* x_ref, const double* x_lin, const double* u_lin,
                                     double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                                     double* slack_opt, int32_t* iters, int8_t* workingSetB, int8_t* workingSetC) {
    int NX, NU, NS;
    if (!p || p->ctx.empty() || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || N < 1) return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    const int W = (int)p->ctx.size();
    const size_t nU = (size_t)NU * N, nV = nU + NS, nXN = (size_t)NX * N;
    const size_t nC = (size_t)((model == FSAE_MODEL_KINEMATIC) ? 6 : 20) * N;
    std::vector<int> rcs(W, FSAE_OK);
    auto shard = [&](int r) {
        int64_t lo, hi;
        fsae_shard_range(B, r, W, &lo, &hi);
        if (hi <= lo) return;
        rcs[r] = fsae_ltvmpc_host(p->ctx[r], model, (int)(hi - lo), N, dt, track_id ? track_id + lo : nullptr,
                                  param_id ? param_id + lo : nullptr, x0 + lo * NX, x_ref + lo * nXN, x_lin + lo * nXN,
                                  u_lin + lo * nU, u_opt + lo * nU, x_opt + lo * nXN, exitflag + lo, fval + lo,
                                  slack_opt + lo * NS, iters ? iters + lo : nullptr,
                                  workingSetB ? workingSetB + lo * nV : nullptr, workingSetC ? workingSetC + lo * nC : nullptr);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < W; ++r) th.emplace_back(shard, r);
    shard(0);                                            // the caller's thread drives the first device itself
    for (auto& t : th) t.join();
    for (int r = 0; r < W; ++r)
        if (rcs[r] != FSAE_OK) { p->err = "device " + std::to_string(p->ctx[r]->device) + ": " + fsae_last_error(p->ctx[r]); return rcs[r]; }
    return FSAE_OK;
}

// merge exit flags / iteration counts across SQP passes
__global__ void sqp_merge_kernel(int B, const int32_t* ef_pass, const int32_t* it_pass, int32_t* ef_acc, int32_t* it_acc, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (first) { ef_acc[i] = ef_pass[i]; it_acc[i] = it_pass[i]; }
    else { if (ef_acc[i] == 0) ef_acc[i] = ef_pass[i]; it_acc[i] += it_pass[i]; }
}

extern "C" int fsae_ltvmpc_sqp_host(fsae_ctx* ctx, int model, int B, int N, double dt, int n_sqp,
                                    const int32_t* track_id, const int32_t* param_id,
                                    const double* x0, const double* x_ref,
                                    const double* x_lin, const double* u_lin,
                                    double* u_opt, double* x_opt, int32_t* exitflag, double* fval,
                                    double* slack_opt, int32_t* iters) {
    int NX, NU, NS;
    if (!ctx || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || n_sqp < 1 || !x0 || !x_ref || !x_lin || !u_lin ||
        !u_opt || !x_opt || !exitflag || !fval || !slack_opt)
        return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    int rc = check_ids(ctx, B, track_id, param_id);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const int nU = NU * N, nXN = NX * N;
    const size_t sz_in[4] = {(size_t)B * NX * 8, (size_t)B * nXN * 8, (size_t)B * nXN * 8, (size_t)B * nU * 8};
    const double* src[4] = {x0, x_ref, x_lin, u_lin};
    for (int i = 0; i < 4; ++i) {
        CK(ctx->in[i].reserve(sz_in[i]));
        CK(cudaMemcpyAsync(ctx->in[i].p, src[i], sz_in[i], cudaMemcpyHostToDevice, ctx->stream));
    }
    const int32_t *d_tid, *d_pid;
    rc = upload_ids(ctx, B, track_id, param_id, &d_tid, &d_pid);
    if (rc) return rc;
    // out[0]/out[1] double as the next pass's u_lin/x_lin (x_opt's stacked layout IS the
    // [N_x x N] column-major layout of x_lin); ping-pong with out[8]/out[9]
    const size_t sz_out[6] = {(size_t)B * nU * 8, (size_t)B * nXN * 8, (size_t)B * 4, (size_t)B * 8, (size_t)B * NS * 8, (size_t)B * 4};
    for (int i = 0; i < 6; ++i) CK(ctx->out[i].reserve(sz_out[i]));
    CK(ctx->out[8].reserve(sz_out[0]));
    CK(ctx->out[9].reserve(sz_out[1]));
    CK(ctx->out[10].reserve(sz_out[2]));
    CK(ctx->out[11].reserve(sz_out[5]));
    const double* xl = (const double*)ctx->in[2].p;
    const double* ul = (const double*)ctx->in[3].p;
    double *uo = nullptr, *xo = nullptr;
    for (int it = 0; it < n_sqp; ++it) {
        uo = (double*)((it & 1) ? ctx->out[8].p : ctx->out[0].p);
        xo = (double*)((it & 1) ? ctx->out[9].p : ctx->out[1].p);
        rc = fsae_ltvmpc_dev(ctx, model, B, N, dt, d_tid, d_pid, (const double*)ctx->in[0].p, (const double*)ctx->in[1].p,
                             xl, ul, uo, xo, (int32_t*)ctx->out[10].p, (double*)ctx->out[3].p, (double*)ctx->out[4].p,
                             (int32_t*)ctx->out[11].p, nullptr, nullptr, ctx->stream);
        if (rc) return rc;
        sqp_merge_kernel<<<(B + 255) / 256, 256, 0, ctx->stream>>>(B, (const int32_t*)ctx->out[10].p, (const int32_t*)ctx->out[11].p,
                                                                  (int32_t*)ctx->out[2].p, (int32_t*)ctx->out[5].p, it == 0);
        ctx->launches++;
        CK(cudaGetLastError());
        xl = xo;
        ul = uo;
    }
    CK(cudaMemcpyAsync(u_opt, uo, sz_out[0], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(x_opt, xo, sz_out[1], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(exitflag, ctx->out[2].p, sz_out[2], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(fval, ctx->out[3].p, sz_out[3], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(slack_opt, ctx->out[4].p, sz_out[4], cudaMemcpyDeviceToHost, ctx->stream));
    if (iters) CK(cudaMemcpyAsync(iters, ctx->out[5].p, sz_out[5], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

extern "C" int fsae_qpoases_host(fsae_ctx* ctx, int B, int nV, int nC,
                                 const double* H, const double* g, const double* A,
                                 const double* lb, const double* ub, const double* lbA, const double* ubA,
                                 double* x, double* fval, int32_t* exitflag, int32_t* iters,
                                 double* lambda, int8_t* workingSetB, int8_t* workingSetC) {
    constexpr int NVMAX = 95, NVBIG = 191;      // two instantiations of the dense kernel (dense_qp.cuh)
    if (!ctx || B < 0 || nV <= 0 || nC < 0 || !H || !g || !lb || !ub || !x || !fval || !exitflag ||
        (nC > 0 && (!A || !lbA || !ubA)))
        return FSAE_ERR_ARG;
    if (nV > NVBIG) { ctx->err = "fsae_qpoases_host: nV > 191 is not supported by this build"; return FSAE_ERR_UNSUPPORTED; }
    if (B == 0) return FSAE_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t szi[7] = {(size_t)B * nV * nV * 8, (size_t)B * nV * 8, (size_t)B * nC * nV * 8, (size_t)B * nV * 8,
                           (size_t)B * nV * 8, (size_t)B * nC * 8, (size_t)B * nC * 8};
    const double* src[7] = {H, g, A, lb, ub, lbA, ubA};
    for (int i = 0; i < 7; ++i) {
        CK(ctx->in[i].reserve(szi[i] ? szi[i] : 8));
        if (szi[i]) CK(cudaMemcpyAsync(ctx->in[i].p, src[i], szi[i], cudaMemcpyHostToDevice, ctx->stream));
    }
    const size_t szo[7] = {(size_t)B * nV * 8, (size_t)B * 8, (size_t)B * 4, (size_t)B * 4, (size_t)B * (nV + nC) * 8,
                           (size_t)B * nV, (size_t)B * (nC ? nC : 1)};
    void* dst[7] = {x, fval, exitflag, iters, lambda, workingSetB, workingSetC};
    for (int i = 0; i < 7; ++i) CK(ctx->out[i].reserve(szo[i]));
    DenseArgs a;
    a.B = B; a.nV = nV; a.nC = nC;
    a.H = (const double*)ctx->in[0].p; a.g = (const double*)ctx->in[1].p; a.A = (const double*)ctx->in[2].p;
    a.lb = (const double*)ctx->in[3].p; a.ub = (const double*)ctx->in[4].p;
    a.lbA = (const double*)ctx->in[5].p; a.ubA = (const double*)ctx->in[6].p;
    a.x = (double*)ctx->out[0].p; a.fval = (double*)ctx->out[1].p; a.exitflag = (int32_t*)ctx->out[2].p;
    a.iters = iters ? (int32_t*)ctx->out[3].p : nullptr;
    a.lambda = lambda ? (double*)ctx->out[4].p : nullptr;
    a.wsB = workingSetB ? (int8_t*)ctx->out[5].p : nullptr;
    a.wsC = (workingSetC && nC) ? (int8_t*)ctx->out[6].p : nullptr;
    a.feas_tol = ctx->h_params[0].feas_tol; a.flat_eps = ctx->h_params[0].flat_eps; a.max_iter = ctx->h_params[0].max_iter;
    a.counters = ctx->d_counters;
    a.hscratch = nullptr;
    if (nV <= NVMAX) {
        const size_t smem = sizeof(DenseSm<NVMAX>) + (size_t)nV + nC + 16;
        auto kern = dense_qp_kernel<NVMAX>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<B, 256, smem, ctx->stream>>>(a);
        ctx->launches++;
        CK(cudaGetLastError());
    } else {
        // 96 <= nV <= 191 (the condensed QPs of horizon 80): 12 warps, half of the operator tile in shared memory, H in the
        // solver's variable order in a global slab (the pool of the long-horizon fused kernels), launched in slices
        using DS = DenseSm<NVBIG, 12, 3, true>;
        const size_t smem = sizeof(DS) + (size_t)nV + nC + 16;
        if (smem > 232448) { ctx->err = "fsae_qpoases_host: too many rows for the large instantiation (shared memory)"; return FSAE_ERR_UNSUPPORTED; }
        auto kern = dense_qp_kernel<NVBIG, 12, 3, true>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        constexpr int SLICE = 1024;
        DevBuf& pool = slab_pool(ctx, ctx->stream);
        const int nsl = B < SLICE ? B : SLICE;
        CK(pool.reserve((size_t)nsl * nV * nV * sizeof(double)));
        for (int lo = 0; lo < B; lo += SLICE) {
            DenseArgs c = a;
            c.B = (lo + SLICE <= B) ? SLICE : B - lo;
            c.hscratch = (double*)pool.p;
            c.H = a.H + (size_t)lo * nV * nV; c.g = a.g + (size_t)lo * nV; c.A = a.A + (size_t)lo * nC * nV;
            c.lb = a.lb + (size_t)lo * nV; c.ub = a.ub + (size_t)lo * nV; c.lbA = a.lbA + (size_t)lo * nC; c.ubA = a.ubA + (size_t)lo * nC;
            c.x = a.x + (size_t)lo * nV; c.fval = a.fval + lo; c.exitflag = a.exitflag + lo;
            if (a.iters) c.iters = a.iters + lo;
            if (a.lambda) c.lambda = a.lambda + (size_t)lo * (nV + nC);
            if (a.wsB) c.wsB = a.wsB + (size_t)lo * nV;
            if (a.wsC) c.wsC = a.wsC + (size_t)lo * nC;
            kern<<<c.B, 384, smem, ctx->stream>>>(c);
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    for (int i = 0; i < 7; ++i)
        if (dst[i] && (i != 6 || nC)) CK(cudaMemcpyAsync(dst[i], ctx->out[i].p, szo[i], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSAE_OK;
}

extern "C" int fsae_closed_loop_host(fsae_ctx* ctx, int model, int B, int N, double dt, int n_sim,
                                     double target_vel, const int32_t* track_id, const int32_t* param_id,
                                     const double* plant0, const double* x_opt0, const double* u_opt0,
                                     double* plant_final, int32_t* steps,
                                     double* n_hist, double* plant_hist, int32_t* exit_hist) {
    int NX, NU, NS;
    if (!ctx || model_dims(model, NX, NU, NS) != FSAE_OK || B < 0 || n_sim < 1 || !plant0 || !x_opt0 || !u_opt0 ||
        !plant_final || !steps)
        return FSAE_ERR_ARG;
    if (B == 0) return FSAE_OK;
    int rc = check_ids(ctx, B, track_id, param_id);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int nU = NU * N, nXN = NX * N;
    const int32_t *d_tid, *d_pid;
    rc = upload_ids(ctx, B, track_id, param_id, &d_tid, &d_pid);
    if (rc) return rc;
    // in[0] x0, in[1] x_ref, in[2]/out[9] x_opt ping-pong, in[3]/out[8] u_opt ping-pong
    const size_t sz_x0 = (size_t)B * NX * 8, sz_x = (size_t)B * nXN * 8, sz_u = (size_t)B * nU * 8;
    CK(ctx->in[0].reserve(sz_x0)); CK(ctx->in[1].reserve(sz_x)); CK(ctx->in[2].reserve(sz_x)); CK(ctx->in[3].reserve(sz_u));
    CK(ctx->out[0].reserve(sz_u)); CK(ctx->out[1].reserve(sz_x)); CK(ctx->out[2].reserve((size_t)B * 4));
    CK(ctx->out[3].reserve((size_t)B * 8)); CK(ctx->out[4].reserve((size_t)B * NS * 8)); CK(ctx->out[5].reserve((size_t)B * 4));
    DevBuf plant, pid, alive, stepsb, nh, ph, eh;
    auto cleanup = [&]() { plant.release(); pid.release(); alive.release(); stepsb.release(); nh.release(); ph.release(); eh.release(); };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return FSAE_ERR_CUDA; } } while (0)
    CKC(plant.reserve((size_t)B * 7 * 8)); CKC(pid.reserve((size_t)B * 4 * 8)); CKC(alive.reserve((size_t)B * 4)); CKC(stepsb.reserve((size_t)B * 4));
    if (n_hist) CKC(nh.reserve((size_t)B * n_sim * 8));
    if (plant_hist) CKC(ph.reserve((size_t)B * n_sim * 7 * 8));
    if (exit_hist) CKC(eh.reserve((size_t)B * n_sim * 4));
    CKC(cudaMemcpyAsync(plant.p, plant0, (size_t)B * 7 * 8, cudaMemcpyHostToDevice, st));
    CKC(cudaMemcpyAsync(ctx->in[2].p, x_opt0, sz_x, cudaMemcpyHostToDevice, st));
    CKC(cudaMemcpyAsync(ctx->in[3].p, u_opt0, sz_u, cudaMemcpyHostToDevice, st));
    CKC(cudaMemsetAsync(pid.p, 0, (size_t)B * 4 * 8, st));
    CKC(cudaMemsetAsync(stepsb.p, 0, (size_t)B * 4, st));
    if (n_hist) CKC(cudaMemsetAsync(nh.p, 0, (size_t)B * n_sim * 8, st));
    if (plant_hist) CKC(cudaMemsetAsync(ph.p, 0, (size_t)B * n_sim * 7 * 8, st));
    if (exit_hist) CKC(cudaMemsetAsync(eh.p, 0, (size_t)B * n_sim * 4, st));
    {
        std::vector<int32_t> ones(B, 1);
        CKC(cudaMemcpyAsync(alive.p, ones.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));
        CKC(cudaStreamSynchronize(st));
    }
    SimArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.N = N; a.model = model; a.dt = dt; a.target_vel = target_vel; a.ramp = 10.0;
    a.track_id = d_tid; a.param_id = d_pid; a.tracks = ctx->d_tracks; a.params = ctx->d_params;
    a.plant = (double*)plant.p; a.pid = (double*)pid.p; a.alive = (int32_t*)alive.p;
    a.x0 = (double*)ctx->in[0].p; a.x_ref = (double*)ctx->in[1].p;
    a.n_hist = n_hist ? (double*)nh.p : nullptr; a.plant_hist = plant_hist ? (double*)ph.p : nullptr;
    a.exit_hist = exit_hist ? (int32_t*)eh.p : nullptr; a.steps = (int32_t*)stepsb.p;
    for (int i = 0; i < FSAE_MAX_TRACKS; ++i) a.track_len[i] = ctx->track_len[i];
    a.n_sim = n_sim;
    const double* xl = (const double*)ctx->in[2].p;
    const double* ul = (const double*)ctx->in[3].p;
    const unsigned grid = (unsigned)((B + 127) / 128);
    for (int step = 0; step <= n_sim; ++step) {
        a.step = step; a.do_plant = step > 0; a.x_opt = xl; a.exitflag = step > 0 ? (const int32_t*)ctx->out[2].p : nullptr;
        sim_advance_kernel<<<grid, 128, 0, st>>>(a);
        ctx->launches++;
        CKC(cudaGetLastError());
        if (step == n_sim) break;
        double* uo = (double*)((step & 1) ? ctx->out[8].p : ctx->out[0].p);
        double* xo = (double*)((step & 1) ? ctx->out[9].p : ctx->out[1].p);
        if (step == 1) { CKC(ctx->out[8].reserve(sz_u)); CKC(ctx->out[9].reserve(sz_x)); uo = (double*)ctx->out[8].p; xo = (double*)ctx->out[9].p; }
        rc = fsae_ltvmpc_dev(ctx, model, B, N, dt, d_tid, d_pid, (const double*)ctx->in[0].p, (const double*)ctx->in[1].p,
                             xl, ul, uo, xo, (int32_t*)ctx->out[2].p, (double*)ctx->out[3].p, (double*)ctx->out[4].p,
                             (int32_t*)ctx->out[5].p, nullptr, nullptr, st);
        if (rc) { cleanup(); return rc; }
        xl = xo;
        ul = uo;
    }
    CKC(cudaMemcpyAsync(plant_final, plant.p, (size_t)B * 7 * 8, cudaMemcpyDeviceToHost, st));
    CKC(cudaMemcpyAsync(steps, stepsb.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (n_hist) CKC(cudaMemcpyAsync(n_hist, nh.p, (size_t)B * n_sim * 8, cudaMemcpyDeviceToHost, st));
    if (plant_hist) CKC(cudaMemcpyAsync(plant_hist, ph.p, (size_t)B * n_sim * 7 * 8, cudaMemcpyDeviceToHost, st));
    if (exit_hist) CKC(cudaMemcpyAsync(exit_hist, eh.p, (size_t)B * n_sim * 4, cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
#undef CKC
    cleanup();
    return FSAE_OK;
}

// debug / test taps -------------------------------------------------------------------
extern "C" int fsae_debug_counters(fsae_ctx* ctx, uint64_t* out3, int reset) {
    if (!ctx || !out3) return FSAE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[3];
    CK(cudaMemcpy(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 3; ++i) out3[i] = h[i];
    if (reset) CK(cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)));
    return FSAE_OK;
}

// measured FP64 FMA peak (TFLOP/s, FMA = 2 flops) of the device: roofline denominator
extern "C" int fsae_probe_fp64_tflops(fsae_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return FSAE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    constexpr int ILP = 8;
    const int iters = 1 << 14;
    const int blocks = prop.multiProcessorCount * 8;
    CK(ctx->out[0].reserve((size_t)blocks * 256 * 8));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        dfma_probe_kernel<ILP><<<blocks, 256, 0, ctx->stream>>>((double*)ctx->out[0].p, iters, 1.0000001, 1e-9);
        ctx->launches++;
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = 2.0 * ILP * (double)iters * blocks * 256.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    ctx->ev_valid = false;
    *tflops = best;
    return FSAE_OK;
}

// Debug taps of the fused kernel (tests): device buffers that the NEXT fsae_ltvmpc_dev calls fill with the
// condensed Hessian H [nV x nV x B], the gradient g [nV x B] and the initial dual active-set operator
// M = [e_slack | J] [nV x nV x B] (column-major per problem).  Null pointers switch a tap off.
extern "C" int fsae_debug_set_taps(fsae_ctx* ctx, double* d_H, double* d_g, double* d_M) {
    if (!ctx) return FSAE_ERR_ARG;
    ctx->tap_H = d_H; ctx->tap_g = d_g; ctx->tap_M = d_M;
    return FSAE_OK;
}

// select the fused kernel variant (tests cross-check v1 against v2); returns the previous one
extern "C" int fsae_debug_set_kernel_version(fsae_ctx* ctx, int v) {
#ifdef FSAE_XCHECK
    if (!ctx || (v != 1 && v != 2 && v != 21 && v != 22 && v != 23 && v != 24 && v != 25 && v != 26 && v != 28 && v != 29 && v != 31 && v != 32)) return FSAE_ERR_ARG;
#else
    if (!ctx || v != 2) return FSAE_ERR_ARG;     // the product library carries the product kernel only
#endif
    const int old = ctx->kernel_version;
    ctx->kernel_version = v;
    return old;
}

#ifdef FSAE_PROFILE
extern "C" int fsae_profile_read(fsae_ctx* ctx, uint64_t* out16, int reset) {
    if (!ctx || !out16) return FSAE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    unsigned long long h[16];
    CK(cudaMemcpyFromSymbol(h, fsae::g_phase_cycles, sizeof(h)));
    for (int i = 0; i < 16; ++i) out16[i] = h[i];
    if (reset) { memset(h, 0, sizeof(h)); CK(cudaMemcpyToSymbol(fsae::g_phase_cycles, h, sizeof(h))); }
    return FSAE_OK;
}
extern "C" int fsae_profile_read_stages(fsae_ctx* ctx, uint64_t* out16, int reset) {
    if (!ctx || !out16) return FSAE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    unsigned long long h[16];
    CK(cudaMemcpyFromSymbol(h, fsae::g_stage_cycles, sizeof(h)));
    for (int i = 0; i < 16; ++i) out16[i] = h[i];
    if (reset) { memset(h, 0, sizeof(h)); CK(cudaMemcpyToSymbol(fsae::g_stage_cycles, h, sizeof(h))); }
    return FSAE_OK;
}
#endif
