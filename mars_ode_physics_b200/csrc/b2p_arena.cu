// b2p_arena.cu -- host side of the C ABI declared in include/mars_ode_b200.h.
//
// An arena is a scene template (bodies, joints, geoms) replicated over W worlds.  Per-world data
// lives in SoA device buffers (layout in b2p_dev.h); small host mirrors with dirty flags make the
// one-world-at-a-time accessors of the reference API (Frame::getPosition, Joint::setVelocity ...)
// work without the caller managing transfers.  Creation-time joint geometry (anchors/axes in body
// frames, relative rotations) is evaluated here in double precision from the template poses, as
// ODE does inside dJointSet*Anchor / dJointSet*Axis.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/mars_ode_b200.h"
#include "b2p_dev.h"
#include "b2p_universal.h"

extern "C" size_t b2p_assemble_smem_bytes(int nb, int nj, int maxc, int islands);
extern "C" size_t b2p_sor_smem_bytes(int nb, int maxrows, int rows_resident);
extern "C" int b2p_sor_register_rows(int maxrows, int wanted);
extern "C" size_t b2p_sor_tmem_smem_bytes(int nb, int maxrows, int pool, int warps);
extern "C" int b2p_sor_tmem_rows_on_chip(int kr, int warps);
extern "C" int b2p_dense_fits_warp_kernel(int nb, int maxrows);
extern "C" void b2p_dense_scratch_sizes(int nb, int maxrows, size_t *mat_doubles, size_t *vec_doubles, size_t *set_bytes);
extern "C" cudaError_t b2p_launch_dense(const StepParams *P, const DenseScratch *D, int nb, int maxrows, int sm_count,
                                        int force_thread_kernel, cudaStream_t stream);
extern "C" cudaError_t b2p_launch_sor_tmem(const StepParams *P, int nb, int maxrows, cudaStream_t stream);
extern "C" cudaError_t b2p_launch_assemble(const StepParams *P, int nb, int nj, int maxc, cudaStream_t stream);
extern "C" cudaError_t b2p_launch_sor(const StepParams *P, int nb, int maxrows, cudaStream_t stream);
extern "C" cudaError_t b2p_launch_integrate(const StepParams *P, int nb, int nj, cudaStream_t stream);
extern "C" cudaError_t b2p_launch_residual(const StepParams *P, float *out, cudaStream_t stream);
extern "C" cudaError_t b2p_launch_collide(const CollideParams *P, int ng, int sampled, cudaStream_t stream);

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) return fail(B2P_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

const double kInf = std::numeric_limits<double>::infinity();

// ---------------- double-precision helpers for creation-time geometry ----------------
struct D3 {
    double v[3];
};
void q_to_R(const double *q, double *R)
{
    double qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3;
    R[1] = 2 * (q[1] * q[2] - q[0] * q[3]);
    R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[3] = 2 * (q[1] * q[2] + q[0] * q[3]);
    R[4] = 1 - qq1 - qq3;
    R[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[6] = 2 * (q[1] * q[3] - q[0] * q[2]);
    R[7] = 2 * (q[2] * q[3] + q[0] * q[1]);
    R[8] = 1 - qq1 - qq2;
}
void Rv(double *r, const double *R, const double *v)
{
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    r[0] = x, r[1] = y, r[2] = z;
}
void Rtv(double *r, const double *R, const double *v)
{
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    r[0] = x, r[1] = y, r[2] = z;
}
double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
void cross(double *r, const double *a, const double *b)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x, r[1] = y, r[2] = z;
}
void norm3(double *a)
{
    double l = dot(a, a);
    if (l > 0) {
        l = 1 / std::sqrt(l);
        a[0] *= l, a[1] *= l, a[2] *= l;
    } else {
        a[0] = 1, a[1] = 0, a[2] = 0;
    }
}
void norm4(double *a)
{
    double l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0) {
        l = 1 / std::sqrt(l);
        for (int i = 0; i < 4; i++) a[i] *= l;
    } else {
        a[0] = 1, a[1] = a[2] = a[3] = 0;
    }
}
void qmul1(double *r, const double *b, const double *c) // b^-1 c
{
    double t[4];
    t[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    t[1] = b[0] * c[1] - b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    t[2] = b[0] * c[2] - b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    t[3] = b[0] * c[3] - b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
    memcpy(r, t, sizeof(t));
}
void qmul2(double *r, const double *b, const double *c) // b c^-1
{
    double t[4];
    t[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    t[1] = -b[0] * c[1] + b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    t[2] = -b[0] * c[2] + b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    t[3] = -b[0] * c[3] + b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
    memcpy(r, t, sizeof(t));
}
void qmul3(double *r, const double *b, const double *c) // b^-1 c^-1
{
    double t[4];
    t[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
    t[1] = -b[0] * c[1] - b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
    t[2] = -b[0] * c[2] - b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
    t[3] = -b[0] * c[3] - b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
    memcpy(r, t, sizeof(t));
}
bool invert33(double *dst, const double *m)
{
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0) return false;
    double r = 1 / det, t[9];
    t[0] = c00 * r;
    t[1] = (m[2] * m[7] - m[1] * m[8]) * r;
    t[2] = (m[1] * m[5] - m[2] * m[4]) * r;
    t[3] = c01 * r;
    t[4] = (m[0] * m[8] - m[2] * m[6]) * r;
    t[5] = (m[2] * m[3] - m[0] * m[5]) * r;
    t[6] = c02 * r;
    t[7] = (m[1] * m[6] - m[0] * m[7]) * r;
    t[8] = (m[0] * m[4] - m[1] * m[3]) * r;
    memcpy(dst, t, sizeof(t));
    return true;
}

struct HLimot {
    double lostop = -kInf, histop = kInf, fudge = 1, normal_cfm = 0, stop_erp = 0, stop_cfm = 0, bounce = 0;
    double vel = 0, fmax = 0; // template value of the per-world commands
};

struct HBody {
    double mass = 1, I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, invI[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int gyroscopic = 1, nogravity = 0, has_max_ang = 0;
    double max_angular_speed = kInf;
    double pos[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0}; // template pose (world 0 / broadcast)
    double lin_damp = 1.0, ang_damp = 1.0;          // Frame::update damping factors; >= 1: off
};

struct HJoint {
    int type = 0, b[2] = {-1, -1}, reverse = 0;
    double anchor1[3] = {0, 0, 0}, anchor2[3] = {0, 0, 0}, axis1[3] = {1, 0, 0}, axis2[3] = {1, 0, 0};
    double qrel[4] = {1, 0, 0, 0}, offset[3] = {0, 0, 0};
    double qrel2[4] = {1, 0, 0, 0}; // universal: body 2 against the cross
    HLimot limot1, limot2;
    double c0 = 0, s0 = 0, v1[3] = {1, 0, 0}, v2[3] = {0, 1, 0}, w1[3] = {1, 0, 0}, w2[3] = {0, 1, 0};
    double susp_erp = 0, susp_cfm = 0, erp = 0, cfm = 0;
    double friction = -0.1; // Joint::applyJointFriction coefficient (Joint.cpp:43: off by default)
};

struct HGeom {
    int cls = 0, body = -1;
    double dims[4] = {0, 0, 0, 0}, lpos[3] = {0, 0, 0}, lquat[4] = {1, 0, 0, 0};
    b2p_surface s{};
};

// A per-world SoA buffer with a host mirror and coherence flags.
template <typename T> struct Mirror {
    std::vector<T> host;
    T *dev = nullptr;
    size_t dev_elems = 0;
    bool host_newer = false, dev_newer = false;
    size_t fields = 0;
};

} // namespace

struct b2p_arena {
    int W = 0, Wp = 0, device = 0;
    cudaStream_t stream = nullptr;
    double gravity[3] = {0, 0, 0}, erp = 0.2, cfm = 1e-5, max_ang = kInf, contact_max_vel = kInf, min_depth = 0,
           sor_w = 1.3;
    int iters = 20, reorder = B2P_REORDER_RANDOMLY;
    std::vector<HBody> bodies;
    std::vector<HJoint> joints;
    std::vector<HGeom> geoms;
    int maxc = 8, rolling = 0;
    bool tmpl_dirty = true;
    DevTemplate *dT = nullptr;
    std::vector<float> hf_heights;
    float *d_hf = nullptr;
    int hf_nx = 0, hf_ny = 0;
    double hf_dx = 0, hf_x0 = 0, hf_y0 = 0;
    bool has_hf = false;
    b2p_surface hf_surface{};

    Mirror<float> state, force, jcmd, jobs, jfb, crec, cfb;
    Mirror<float> bodyx, jupd; // post-step outputs (b2p_dev.h); sized when post_flags selects them
    int post_flags = B2P_POST_JOINT_UPDATE | B2P_POST_FRAME_UPDATE;
    Mirror<int> ccount, nrows;
    Mirror<uint32_t> seed;
    Mirror<uint8_t> jrows;
    // solver scratch (device only)
    float4 *d_rows = nullptr;
    float *d_lambda = nullptr, *d_rowS = nullptr, *d_invIw = nullptr, *d_fcacc = nullptr;
    uint32_t *d_cmap = nullptr, *d_isl = nullptr;
    uint8_t *d_ord0 = nullptr;
    int order_mode = B2P_ORDER_WORLD_BLOCK;
    size_t scratch_rows = 0, scratch_nb = 0, scratch_maxc = 0;
    int rows_resident = 0;   // SOR kernel: sweep positions kept in shared memory; 0 = automatic
    int rows_registers = -1; // SOR kernel: sweep positions kept in registers; < 0 = automatic
    int sor_kernel = -1;     // B2P_SOR_AUTO / B2P_SOR_PAIRED (b2p_sor.cu) / B2P_SOR_TMEM (b2p_sor_tmem.cu)
    int sor_warps = 0;       // Tensor-Memory kernel: warps per CTA (4 or 8); 0 = automatic
    int sm_count = 0;
    int sor_kernel_used = 0;
    int stepper = 0;      // B2P_STEPPER_QUICK / B2P_STEPPER_DENSE
    DenseScratch dense{}; // scratch of the dense stepper (allocated on first use)
    int dense_rows = 0, dense_nb = 0;
    int tune = -1; // experiment switches (b2p_arena_set_tune); < 0: from the environment variable B2P_TUNE
    int *d_hw_rows = nullptr, *h_hw_rows = nullptr; // largest row count of any world so far (device / pinned copy)
    // optional per-kernel timing (bench / profiling): events around the three launches of a step
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double kernel_ms[3] = {0, 0, 0};
    long long timed_steps = 0;
    int *d_pairs = nullptr, *d_npairs = nullptr;
    int maxpairs = 0;
    int launches = 0;
    bool want_pairs = false; // the broadphase pair list is a debug output: written only after it was asked for
    int *d_dropped = nullptr; // contacts the narrowphase found beyond max_contacts, summed over worlds and steps
    float *d_residual = nullptr;
    StepParams lastP{};      // parameters of the last step (b2p_debug_lcp_residual)
    // one CUDA graph per launch shape: [0] = step, [1] = collide + step.  Re-captured when any kernel parameter
    // changes (step size, gravity, on-chip row split ...), launched with one call otherwise.
    struct StepGraph {
        // [0] (collide +) assemble, [1] SOR + integrate: the event the command uploads wait for is recorded on the
        // stream between the two (recorded from inside one graph it only becomes visible when that graph ends)
        // [2] all of it in one graph, used as long as nobody uploads per-tick commands
        cudaGraphExec_t exec[3] = {nullptr, nullptr, nullptr};
        StepParams P{};
        CollideParams C{};
        int use_tmem = -1, sampled = -1;
    } graphs[2];
    bool use_graph = true;
    bool upload_mode = false; // per-tick command uploads are in use (b2p_joints_set_param_batch): split step graph
    cudaStream_t cap_stream = nullptr; // capture happens here (the arena's own stream may be the legacy default stream)
    // pipelined observation read-back (b2p_body_get_state_batch_begin / b2p_io_wait): a copy stream and two
    // staging slots, so that the D2H of step k overlaps the kernels of step k+1
    struct IoSlot {
        float *dev = nullptr;
        size_t elems = 0;
        float *dev_jobs = nullptr; // joint observations snapshot (b2p_observe_begin)
        size_t jobs_elems = 0;
        cudaEvent_t snap = nullptr, done = nullptr;
        bool pending = false;
    } io[2];
    cudaStream_t io_stream = nullptr;
    cudaStream_t up_stream = nullptr; // command uploads: their own stream, so that they never queue behind a read-back
    int io_next = 0;
    // command uploads on the copy stream: ordered after the assemble kernel of the last step (the only reader of
    // jcmd) and before the next one
    cudaEvent_t ev_asm_done = nullptr, ev_h2d_done = nullptr;
    bool asm_recorded = false, h2d_pending = false;
    long long steps_done = 0;
    bool stepped = false;
};

namespace {

template <typename T> int mirror_resize(b2p_arena *a, Mirror<T> &m, size_t fields, T fill)
{
    const size_t n = fields * (size_t)a->Wp;
    if (m.fields == fields && m.host.size() == n) return 0;
    // device content (if newer) must come home before the host vector is regrown
    if (m.dev_newer && m.dev && m.dev_elems) {
        CUDA_TRY(cudaMemcpyAsync(m.host.data(), m.dev, sizeof(T) * std::min(m.dev_elems, m.host.size()),
                                 cudaMemcpyDeviceToHost, a->stream));
        CUDA_TRY(cudaStreamSynchronize(a->stream));
        m.dev_newer = false;
    }
    m.host.resize(n, fill);
    m.fields = fields;
    m.host_newer = true;
    return 0;
}

template <typename T> int mirror_to_device(b2p_arena *a, Mirror<T> &m)
{
    const size_t n = m.host.size();
    if (m.dev_elems != n) {
        if (m.dev) CUDA_TRY(cudaFree(m.dev));
        m.dev = nullptr;
        m.dev_elems = 0;
        if (n) {
            CUDA_TRY(cudaMalloc(&m.dev, sizeof(T) * n));
            m.dev_elems = n;
        }
        m.host_newer = true;
    }
    if (m.host_newer && n) {
        CUDA_TRY(cudaMemcpyAsync(m.dev, m.host.data(), sizeof(T) * n, cudaMemcpyHostToDevice, a->stream));
        m.host_newer = false;
    }
    return 0;
}

// a command upload may still be in flight on the copy stream: everything on the main stream that follows is
// ordered behind it
int join_pending_upload(b2p_arena *a)
{
    if (a->h2d_pending) {
        CUDA_TRY(cudaStreamWaitEvent(a->stream, a->ev_h2d_done, 0));
        a->h2d_pending = false;
    }
    return 0;
}

template <typename T> int mirror_to_host(b2p_arena *a, Mirror<T> &m)
{
    if (int rc = join_pending_upload(a)) return rc;
    if (m.dev_newer && m.dev && m.dev_elems) {
        CUDA_TRY(cudaMemcpyAsync(m.host.data(), m.dev, sizeof(T) * m.dev_elems, cudaMemcpyDeviceToHost, a->stream));
        CUDA_TRY(cudaStreamSynchronize(a->stream));
    }
    m.dev_newer = false;
    return 0;
}

// One world's column of `n` consecutive fields: device -> host mirror.  The one-world accessors of the reference
// API (Frame::getPosition, Joint::getVelocity ...) read a handful of values; they must not cost a transfer of the
// whole batch.  The mirror stays "device newer": only this column of the host copy is refreshed.
template <typename T> int mirror_fetch_column(b2p_arena *a, Mirror<T> &m, size_t field0, size_t n, int world)
{
    if (!(m.dev_newer && m.dev && m.dev_elems)) return 0;
    if (int rc = join_pending_upload(a)) return rc;
    const size_t pitch = sizeof(T) * (size_t)a->Wp, off = field0 * (size_t)a->Wp + (size_t)world;
    CUDA_TRY(cudaMemcpy2DAsync(m.host.data() + off, pitch, m.dev + off, pitch, sizeof(T), n, cudaMemcpyDeviceToHost,
                               a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

// ... and host mirror -> device, for a one-world setter while the device copy is the newer one
template <typename T> int mirror_push_column(b2p_arena *a, Mirror<T> &m, size_t field0, size_t n, int world)
{
    if (!(m.dev && m.dev_elems)) return 0;
    const size_t pitch = sizeof(T) * (size_t)a->Wp, off = field0 * (size_t)a->Wp + (size_t)world;
    CUDA_TRY(cudaMemcpy2DAsync(m.dev + off, pitch, m.host.data() + off, pitch, sizeof(T), n, cudaMemcpyHostToDevice,
                               a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream)); // the source is pageable host memory that may change right away
    return 0;
}

template <typename T> void mirror_free(Mirror<T> &m)
{
    if (m.dev) cudaFree(m.dev);
    m.dev = nullptr;
}

// set a field for one world or all worlds
template <typename T> void mset(b2p_arena *a, Mirror<T> &m, size_t field, int world, T v)
{
    T *row = m.host.data() + field * (size_t)a->Wp;
    if (world == B2P_ALL_WORLDS) {
        for (int w = 0; w < a->W; w++) row[w] = v;
    } else {
        row[world] = v;
    }
    m.host_newer = true;
}

int check_world(const b2p_arena *a, int world, bool allow_all)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    cudaSetDevice(a->device); // every entry point works on the arena's device, whatever the caller's current one is
    if (world == B2P_ALL_WORLDS && allow_all) return 0;
    if (world < 0 || world >= a->W) return fail(B2P_EINVAL, "world index %d out of range [0,%d)", world, a->W);
    return 0;
}
int check_body(const b2p_arena *a, int body)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    cudaSetDevice(a->device);
    if (body < 0 || body >= (int)a->bodies.size()) return fail(B2P_EINVAL, "body id %d out of range", body);
    return 0;
}
int check_joint(const b2p_arena *a, int j)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    cudaSetDevice(a->device);
    if (j < 0 || j >= (int)a->joints.size()) return fail(B2P_EINVAL, "joint id %d out of range", j);
    return 0;
}

int rpc_of(const b2p_arena *a) { return a->rolling ? 6 : 3; }

// host view of a body's pose in one world (double from the float mirror)
int body_pose(b2p_arena *a, int body, int world, double *pos, double *q)
{
    int rc = mirror_fetch_column(a, a->state, (size_t)body * B2P_STATE_FIELDS, 7, world);
    if (rc) return rc;
    const float *s = a->state.host.data();
    const size_t f = (size_t)body * B2P_STATE_FIELDS;
    for (int i = 0; i < 3; i++) pos[i] = s[(f + i) * a->Wp + world];
    for (int i = 0; i < 4; i++) q[i] = s[(f + 3 + i) * a->Wp + world];
    return 0;
}

int body_vec(b2p_arena *a, Mirror<float> &m, size_t field0, int world, int n, double *out)
{
    int rc = mirror_fetch_column(a, m, field0, (size_t)n, world);
    if (rc) return rc;
    for (int i = 0; i < n; i++) out[i] = m.host[(field0 + i) * a->Wp + world];
    return 0;
}

void fill_limot(DevLimot &d, const HLimot &h)
{
    d.lostop = (float)h.lostop;
    d.histop = (float)h.histop;
    d.fudge_factor = (float)h.fudge;
    d.normal_cfm = (float)h.normal_cfm;
    d.stop_erp = (float)h.stop_erp;
    d.stop_cfm = (float)h.stop_cfm;
    d.bounce = (float)h.bounce;
    d.pad = 0;
}

bool limot_has_stops(const HLimot &l) { return (l.lostop > -kInf || l.histop < kInf) && l.lostop < l.histop; }

void free_scratch(b2p_arena *a)
{
    if (a->d_rows) cudaFree(a->d_rows);
    if (a->d_lambda) cudaFree(a->d_lambda);
    if (a->d_rowS) cudaFree(a->d_rowS);
    if (a->d_invIw) cudaFree(a->d_invIw);
    if (a->d_fcacc) cudaFree(a->d_fcacc);
    if (a->d_cmap) cudaFree(a->d_cmap);
    if (a->d_isl) cudaFree(a->d_isl);
    if (a->d_ord0) cudaFree(a->d_ord0);
    a->d_isl = nullptr;
    a->d_ord0 = nullptr;
    if (a->d_hw_rows) cudaFree(a->d_hw_rows);
    a->d_hw_rows = nullptr;
    a->d_rows = nullptr;
    a->d_lambda = a->d_rowS = a->d_invIw = a->d_fcacc = nullptr;
    a->d_cmap = nullptr;
    a->scratch_rows = a->scratch_nb = a->scratch_maxc = 0;
    if (a->dense.A) cudaFree(a->dense.A);
    if (a->dense.L) cudaFree(a->dense.L);
    if (a->dense.vec) cudaFree(a->dense.vec);
    if (a->dense.sets) cudaFree(a->dense.sets);
    if (a->dense.failures) cudaFree(a->dense.failures);
    if (a->dense.residual) cudaFree(a->dense.residual);
    a->dense = DenseScratch{};
    a->dense_rows = a->dense_nb = 0;
}

// Square root of a symmetric positive (semi)definite 3x3 matrix (row-major) by cyclic Jacobi rotations:
// A = V diag(e) V^T  ->  sqrt(A) = V diag(sqrt(e)) V^T.  Used for the mass-scaled row form M^-1/2.
void sym_sqrt33(const double *A, double *out)
{
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = 0.5 * (A[i * 3 + j] + A[j * 3 + i]);
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; k++) { // A <- A G
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) { // A <- G^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) { // V <- V G
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double x = 0;
            for (int k = 0; k < 3; k++) x += v[i][k] * std::sqrt(std::max(a[k][k], 0.0)) * v[j][k];
            out[i * 3 + j] = x;
        }
}

// (re)build the device template and size every buffer
int finalize(b2p_arena *a)
{
    if (a->jcmd.host_newer)
        if (int rc = join_pending_upload(a)) return rc; // a host-side edit is about to overwrite the device copy
    const int nb = (int)a->bodies.size(), nj = (int)a->joints.size(), ng = (int)a->geoms.size();
    if (nb > B2P_MAX_BODIES) return fail(B2P_ECAPACITY, "too many bodies (%d > %d)", nb, B2P_MAX_BODIES);
    if (nj > B2P_MAX_JOINTS) return fail(B2P_ECAPACITY, "too many joints (%d > %d)", nj, B2P_MAX_JOINTS);
    if (ng > B2P_MAX_GEOMS) return fail(B2P_ECAPACITY, "too many geoms (%d > %d)", ng, B2P_MAX_GEOMS);
    const int rpc = rpc_of(a);
    const int njslots = nj * B2P_ROWS_PER_JOINT, maxrows = njslots + a->maxc * rpc;
    if (maxrows > 253) return fail(B2P_ECAPACITY, "too many constraint rows per world (%d > 253)", maxrows);
    if (a->tmpl_dirty) {
        std::vector<char> buf(sizeof(DevTemplate), 0);
        DevTemplate *T = reinterpret_cast<DevTemplate *>(buf.data());
        T->nb = nb;
        T->nj = nj;
        T->ng = ng;
        T->maxc = a->maxc;
        T->rpc = rpc;
        T->njslots = njslots;
        T->maxrows = maxrows;
        T->has_limit_motors = 0;
        for (int b = 0; b < nb; b++) {
            const HBody &h = a->bodies[b];
            DevBody &d = T->bodies[b];
            d.mass = (float)h.mass;
            d.invMass = (float)(1.0 / h.mass);
            for (int i = 0; i < 9; i++) {
                d.I[i] = (float)h.I[i];
                d.invI[i] = (float)h.invI[i];
            }
            d.sqrtInvMass = (float)std::sqrt(1.0 / h.mass);
            d.sqrtMass = (float)std::sqrt(h.mass);
            {
                double sI[9], sInvI[9];
                sym_sqrt33(h.I, sI);
                sym_sqrt33(h.invI, sInvI);
                for (int i = 0; i < 9; i++) {
                    d.sqrtI[i] = (float)sI[i];
                    d.sqrtInvI[i] = (float)sInvI[i];
                }
            }
            d.gyroscopic = h.gyroscopic;
            d.nogravity = h.nogravity;
            d.has_max_ang = h.has_max_ang;
            d.max_angular_speed = (float)h.max_angular_speed;
            d.lin_damp = (float)h.lin_damp;
            d.ang_damp = (float)h.ang_damp;
        }
        for (int j = 0; j < nj; j++) {
            const HJoint &h = a->joints[j];
            DevJoint &d = T->joints[j];
            d.type = h.type;
            d.b0 = h.b[0];
            d.b1 = h.b[1];
            d.reverse = h.reverse;
            d.slot_base = j * B2P_ROWS_PER_JOINT;
            for (int i = 0; i < 3; i++) {
                d.anchor1[i] = (float)h.anchor1[i];
                d.anchor2[i] = (float)h.anchor2[i];
                d.axis1[i] = (float)h.axis1[i];
                d.axis2[i] = (float)h.axis2[i];
                d.offset[i] = (float)h.offset[i];
                d.v1[i] = (float)h.v1[i];
                d.v2[i] = (float)h.v2[i];
                d.w1[i] = (float)h.w1[i];
                d.w2[i] = (float)h.w2[i];
            }
            for (int i = 0; i < 4; i++) d.qrel[i] = (float)h.qrel[i], d.qrel2[i] = (float)h.qrel2[i];
            fill_limot(d.limot1, h.limot1);
            fill_limot(d.limot2, h.limot2);
            d.c0 = (float)h.c0;
            d.s0 = (float)h.s0;
            d.susp_erp = (float)h.susp_erp;
            d.susp_cfm = (float)h.susp_cfm;
            d.erp = (float)h.erp;
            d.cfm = (float)h.cfm;
            d.friction = (float)h.friction;
            if (h.type != B2P_JOINT_FIXED && limot_has_stops(h.limot1)) T->has_limit_motors = 1;
            if (h.b[0] >= 0 && h.b[1] >= 0) {
                T->connected[h.b[0]] |= 1u << h.b[1];
                T->connected[h.b[1]] |= 1u << h.b[0];
            }
        }
        for (int g = 0; g < ng; g++) {
            const HGeom &h = a->geoms[g];
            DevGeom &d = T->geoms[g];
            d.cls = h.cls;
            d.body = h.body;
            for (int i = 0; i < 4; i++) {
                d.dims[i] = (float)h.dims[i];
                d.lquat[i] = (float)h.lquat[i];
            }
            for (int i = 0; i < 3; i++) d.lpos[i] = (float)h.lpos[i];
            d.has_offset = (h.lpos[0] != 0 || h.lpos[1] != 0 || h.lpos[2] != 0 || h.lquat[0] != 1);
            d.mu = (float)h.s.mu;
            d.mu2 = (float)h.s.mu2;
            d.rho = (float)h.s.rho;
            d.rho2 = (float)h.s.rho2;
            d.rhoN = (float)h.s.rhoN;
            d.soft_erp = (float)h.s.soft_erp;
            d.soft_cfm = (float)h.s.soft_cfm;
        }
        if (a->has_hf) {
            if (!a->d_hf) {
                CUDA_TRY(cudaMalloc(&a->d_hf, sizeof(float) * a->hf_heights.size()));
                CUDA_TRY(cudaMemcpyAsync(a->d_hf, a->hf_heights.data(), sizeof(float) * a->hf_heights.size(),
                                         cudaMemcpyHostToDevice, a->stream));
            }
            T->hf.heights = a->d_hf;
            T->hf.nx = a->hf_nx;
            T->hf.ny = a->hf_ny;
            T->hf.dx = (float)a->hf_dx;
            T->hf.x0 = (float)a->hf_x0;
            T->hf.y0 = (float)a->hf_y0;
            float mn = 1e30f, mx = -1e30f;
            for (float v : a->hf_heights) {
                mn = std::min(mn, v);
                mx = std::max(mx, v);
            }
            T->hf.hmin = mn;
            T->hf.hmax = mx;
        }
        if (!a->dT) CUDA_TRY(cudaMalloc(&a->dT, sizeof(DevTemplate)));
        CUDA_TRY(cudaMemcpyAsync(a->dT, T, sizeof(DevTemplate), cudaMemcpyHostToDevice, a->stream));
        CUDA_TRY(cudaStreamSynchronize(a->stream)); // buf goes out of scope
        a->tmpl_dirty = false;
    }
    int rc;
    if ((rc = mirror_resize(a, a->crec, (size_t)a->maxc * B2P_CONTACT_FIELDS, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->cfb, (size_t)a->maxc * 3, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->ccount, 1, 0))) return rc;
    if ((rc = mirror_resize(a, a->nrows, 1, 0))) return rc;
    if ((rc = mirror_resize(a, a->jobs, (size_t)nj * 4, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->jfb, (size_t)nj * 13, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->jrows, (size_t)nj, (uint8_t)0))) return rc;
    if ((rc = mirror_resize(a, a->bodyx, (a->post_flags & B2P_POST_FRAME_UPDATE) ? (size_t)nb * 15 : 0, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->jupd, (a->post_flags & B2P_POST_JOINT_UPDATE) ? (size_t)nj * 7 : 0, 0.f))) return rc;
    if ((rc = mirror_to_device(a, a->bodyx))) return rc;
    if ((rc = mirror_to_device(a, a->jupd))) return rc;
    if ((rc = mirror_to_device(a, a->state))) return rc;
    if ((rc = mirror_to_device(a, a->force))) return rc;
    if ((rc = mirror_to_device(a, a->jcmd))) return rc;
    if ((rc = mirror_to_device(a, a->crec))) return rc;
    if ((rc = mirror_to_device(a, a->ccount))) return rc;
    if ((rc = mirror_to_device(a, a->seed))) return rc;
    // outputs: make sure device buffers exist (content irrelevant)
    if ((rc = mirror_to_device(a, a->jobs))) return rc;
    if ((rc = mirror_to_device(a, a->jfb))) return rc;
    if ((rc = mirror_to_device(a, a->cfb))) return rc;
    if ((rc = mirror_to_device(a, a->nrows))) return rc;
    if ((rc = mirror_to_device(a, a->jrows))) return rc;
    if (a->scratch_rows != (size_t)maxrows || a->scratch_nb != (size_t)nb || a->scratch_maxc != (size_t)a->maxc ||
        !a->d_ord0) {
        free_scratch(a);
        const size_t mr = (size_t)std::max(maxrows, 1), Wp = (size_t)a->Wp;
        CUDA_TRY(cudaMalloc(&a->d_rows, sizeof(float4) * (mr + 1) * 4 * Wp));
        CUDA_TRY(cudaMalloc(&a->d_lambda, sizeof(float) * mr * Wp));
        CUDA_TRY(cudaMalloc(&a->d_rowS, sizeof(float) * mr * Wp));
        CUDA_TRY(cudaMalloc(&a->d_invIw, sizeof(float) * (size_t)std::max(nb, 1) * 9 * Wp));
        CUDA_TRY(cudaMalloc(&a->d_fcacc, sizeof(float) * (size_t)std::max(nb, 1) * 6 * Wp));
        CUDA_TRY(cudaMalloc(&a->d_cmap, sizeof(uint32_t) * (size_t)std::max(a->maxc, 1) * Wp));
        CUDA_TRY(cudaMalloc(&a->d_isl, sizeof(uint32_t) * (size_t)std::max(nb, 1) * Wp));
        CUDA_TRY(cudaMalloc(&a->d_ord0, mr * Wp));
        CUDA_TRY(cudaMemsetAsync(a->d_isl, 0, sizeof(uint32_t) * (size_t)std::max(nb, 1) * Wp, a->stream));
        CUDA_TRY(cudaMemsetAsync(a->d_ord0, 0, mr * Wp, a->stream));
        CUDA_TRY(cudaMalloc(&a->d_hw_rows, sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(a->d_hw_rows, 0, sizeof(int), a->stream));
        if (!a->h_hw_rows) CUDA_TRY(cudaMallocHost(&a->h_hw_rows, sizeof(int)));
        *a->h_hw_rows = 0;
        // padded worlds (w >= W) and unused slots are read but never used: keep them finite
        CUDA_TRY(cudaMemsetAsync(a->d_rows, 0, sizeof(float4) * (mr + 1) * 4 * Wp, a->stream));
        {
            // slot maxrows: the all-zero row the SOR kernel puts behind the last row of a world.  Its meta names
            // the dummy multiplier (slot maxrows), the constant-1 multiplier (maxrows + 1) and the dummy body nb
            const uint32_t zm = (uint32_t)maxrows | ((uint32_t)(maxrows + 1) << 8) | ((uint32_t)nb << 16) |
                                ((uint32_t)nb << 24);
            std::vector<float4> z(Wp * 2); // unit 1 of slot maxrows: [Wp][2 float4]
            float mf;
            memcpy(&mf, &zm, sizeof(mf));
            for (size_t i = 0; i < Wp; i++) z[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f), z[2 * i + 1] = make_float4(0.f, 0.f, 0.f, mf);
            CUDA_TRY(cudaMemcpyAsync(a->d_rows + ((size_t)maxrows * 2 + 1) * Wp * 2, z.data(), sizeof(float4) * Wp * 2,
                                     cudaMemcpyHostToDevice, a->stream));
            CUDA_TRY(cudaStreamSynchronize(a->stream));
        }
        CUDA_TRY(cudaMemsetAsync(a->d_lambda, 0, sizeof(float) * mr * Wp, a->stream));
        CUDA_TRY(cudaMemsetAsync(a->d_rowS, 0, sizeof(float) * mr * Wp, a->stream)); // (unused slots are read, then masked)
        a->scratch_rows = (size_t)maxrows;
        a->scratch_nb = (size_t)nb;
        a->scratch_maxc = (size_t)a->maxc;
    }
    return 0;
}

// The joint geometry setters below read the poses of the joint's bodies, as ODE's dJointSet*Anchor / *Axis /
// dJointSetFixed read the live dBody.  The template pose (HBody.pos / q) is what they use; once the worlds have
// moved (a step, a state upload, a per-world setter) it is refreshed from world 0 first.  Joint geometry is per
// template: it is shared by all worlds.
int refresh_template_poses(b2p_arena *a, const HJoint &j)
{
    for (int k = 0; k < 2; k++) {
        const int b = j.b[k];
        if (b < 0 || !(a->state.dev_newer && a->state.dev)) continue;
        int rc = mirror_fetch_column(a, a->state, (size_t)b * B2P_STATE_FIELDS, 7, 0);
        if (rc) return rc;
        const float *st = a->state.host.data() + (size_t)b * B2P_STATE_FIELDS * a->Wp;
        HBody &hb = a->bodies[b];
        for (int i = 0; i < 3; i++) hb.pos[i] = st[(size_t)i * a->Wp];
        for (int i = 0; i < 4; i++) hb.q[i] = st[(size_t)(3 + i) * a->Wp];
    }
    return 0;
}

// ---- creation-time joint geometry (ODE joint.cpp setAnchors / setAxes etc.) ----
void set_anchors(b2p_arena *a, HJoint &j, double x, double y, double z)
{
    if (j.b[0] < 0) return;
    const HBody &b0 = a->bodies[j.b[0]];
    double R[9], q[3] = {x - b0.pos[0], y - b0.pos[1], z - b0.pos[2]};
    q_to_R(b0.q, R);
    Rtv(j.anchor1, R, q);
    if (j.b[1] >= 0) {
        const HBody &b1 = a->bodies[j.b[1]];
        double q2[3] = {x - b1.pos[0], y - b1.pos[1], z - b1.pos[2]};
        q_to_R(b1.q, R);
        Rtv(j.anchor2, R, q2);
    } else {
        j.anchor2[0] = x, j.anchor2[1] = y, j.anchor2[2] = z;
    }
}
void set_axes(b2p_arena *a, HJoint &j, double x, double y, double z, double *axis1, double *axis2)
{
    if (j.b[0] < 0) return;
    double q[3] = {x, y, z}, R[9];
    norm3(q);
    if (axis1) {
        q_to_R(a->bodies[j.b[0]].q, R);
        Rtv(axis1, R, q);
    }
    if (axis2) {
        if (j.b[1] >= 0) {
            q_to_R(a->bodies[j.b[1]].q, R);
            Rtv(axis2, R, q);
        } else {
            axis2[0] = q[0], axis2[1] = q[1], axis2[2] = q[2];
        }
    }
}
void initial_relative_rotation(b2p_arena *a, HJoint &j)
{
    if (j.b[0] < 0) return;
    const HBody &b0 = a->bodies[j.b[0]];
    if (j.b[1] >= 0) {
        qmul1(j.qrel, b0.q, a->bodies[j.b[1]].q);
    } else {
        j.qrel[0] = b0.q[0];
        for (int i = 1; i < 4; i++) j.qrel[i] = -b0.q[i];
    }
}
// universal joint: world axes at creation poses, and dxJointUniversal::computeInitialRelativeRotations
void universal_axes_world(const HJoint &j, const double *q0, const double *q1, double *ax1, double *ax2)
{
    double R[9];
    q_to_R(q0, R);
    Rv(ax1, R, j.axis1);
    if (j.b[1] >= 0) {
        q_to_R(q1, R);
        Rv(ax2, R, j.axis2);
    } else {
        ax2[0] = j.axis2[0], ax2[1] = j.axis2[1], ax2[2] = j.axis2[2];
    }
}
void universal_initial_rotations(b2p_arena *a, HJoint &j)
{
    if (j.b[0] < 0) return;
    const double ident[4] = {1, 0, 0, 0};
    const double *q0 = a->bodies[j.b[0]].q, *q1 = j.b[1] >= 0 ? a->bodies[j.b[1]].q : ident;
    double ax1[3], ax2[3];
    universal_axes_world(j, q0, q1, ax1, ax2);
    b2p_uni::initial_rotations<double>(ax1, ax2, q0, q1, j.b[1] >= 0, j.qrel, j.qrel2);
}
void hinge2_axes_world(b2p_arena *a, const HJoint &j, double *ax1, double *ax2)
{
    double R[9];
    q_to_R(a->bodies[j.b[0]].q, R);
    Rv(ax1, R, j.axis1);
    q_to_R(a->bodies[j.b[1]].q, R);
    Rv(ax2, R, j.axis2);
}
void hinge2_axis_info(b2p_arena *a, HJoint &j)
{
    double ax1[3], ax2[3], ax[3];
    hinge2_axes_world(a, j, ax1, ax2);
    cross(ax, ax1, ax2);
    j.s0 = std::sqrt(dot(ax, ax));
    j.c0 = dot(ax1, ax2);
}
bool same3(const double *x, const double *y) { return x[0] == y[0] && x[1] == y[1] && x[2] == y[2]; }
bool zero3(const double *x) { return x[0] == 0 && x[1] == 0 && x[2] == 0; }
void hinge2_make_v(b2p_arena *a, HJoint &j)
{
    if (j.b[0] < 0 || j.b[1] < 0) return;
    double ax1[3], ax2[3], v[3], R[9];
    hinge2_axes_world(a, j, ax1, ax2);
    if (zero3(ax1) || zero3(ax2) || same3(ax1, ax2)) return;
    double k = dot(ax1, ax2);
    for (int i = 0; i < 3; i++) ax2[i] -= k * ax1[i];
    norm3(ax2);
    cross(v, ax1, ax2);
    q_to_R(a->bodies[j.b[0]].q, R);
    Rtv(j.v1, R, ax2);
    Rtv(j.v2, R, v);
}
void hinge2_make_w(b2p_arena *a, HJoint &j)
{
    if (j.b[0] < 0 || j.b[1] < 0) return;
    double ax1[3], ax2[3], v[3], R[9];
    hinge2_axes_world(a, j, ax1, ax2);
    if (zero3(ax1) || zero3(ax2) || same3(ax1, ax2)) return;
    double k = dot(ax2, ax1);
    for (int i = 0; i < 3; i++) ax1[i] -= k * ax2[i];
    norm3(ax1);
    cross(v, ax2, ax1);
    q_to_R(a->bodies[j.b[1]].q, R);
    Rtv(j.w1, R, ax1);
    Rtv(j.w2, R, v);
}

HLimot *limot_for(HJoint &j, int param, int *num)
{
    if ((param & 0xff00) == B2P_PARAM_GROUP2) {
        *num = param & 0xff;
        return &j.limot2;
    }
    *num = param;
    return &j.limot1;
}

// ---- measurement of joint coordinates on the host (getters) ----
double hinge_angle_host(const HJoint &j, const double *q0, const double *q1)
{
    double qrel[4];
    if (j.b[1] >= 0) {
        double qq[4];
        qmul1(qq, q0, q1);
        qmul2(qrel, qq, j.qrel);
    } else {
        qmul3(qrel, q0, j.qrel);
    }
    double cost2 = qrel[0], sint2 = std::sqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    double theta = (dot(qrel + 1, j.axis1) >= 0) ? 2 * std::atan2(sint2, cost2) : 2 * std::atan2(sint2, -cost2);
    if (theta > M_PI) theta -= 2 * M_PI;
    return -theta;
}

} // namespace

// =====================================================================================
extern "C" {

const char *b2p_last_error(void) { return g_err.c_str(); }
int b2p_version(void) { return 1; }

int b2p_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b2p_arena_create(int num_worlds, int device, b2p_arena **out)
{
    if (!out) return fail(B2P_EINVAL, "out is null");
    *out = nullptr;
    if (num_worlds <= 0) return fail(B2P_EINVAL, "num_worlds must be positive");
    int n = b2p_device_count();
    if (n <= 0) return fail(B2P_ENODEVICE, "no CUDA device: the batched step has no CPU fallback");
    if (device < 0 || device >= n) return fail(B2P_EINVAL, "device %d out of range [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    b2p_arena *a = new b2p_arena();
    a->W = num_worlds;
    a->Wp = (num_worlds + 31) & ~31;
    a->device = device;
    cudaDeviceGetAttribute(&a->sm_count, cudaDevAttrMultiProcessorCount, device);
    a->seed.fields = 1;
    a->seed.host.assign((size_t)a->Wp, 0u);
    a->seed.host_newer = true;
    *out = a;
    return 0;
}

void b2p_arena_destroy(b2p_arena *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    cudaStreamSynchronize(a->stream);
    mirror_free(a->state);
    mirror_free(a->force);
    mirror_free(a->jcmd);
    mirror_free(a->jobs);
    mirror_free(a->jfb);
    mirror_free(a->crec);
    mirror_free(a->cfb);
    mirror_free(a->ccount);
    mirror_free(a->nrows);
    mirror_free(a->seed);
    mirror_free(a->jrows);
    mirror_free(a->bodyx);
    mirror_free(a->jupd);
    free_scratch(a);
    for (b2p_arena::StepGraph &g : a->graphs)
        for (int i = 0; i < 3; i++)
            if (g.exec[i]) cudaGraphExecDestroy(g.exec[i]);
    if (a->d_dropped) cudaFree(a->d_dropped);
    if (a->d_residual) cudaFree(a->d_residual);
    for (cudaEvent_t e : a->ev)
        if (e) cudaEventDestroy(e);
    if (a->dT) cudaFree(a->dT);
    if (a->d_hf) cudaFree(a->d_hf);
    if (a->d_pairs) cudaFree(a->d_pairs);
    if (a->d_npairs) cudaFree(a->d_npairs);
    if (a->h_hw_rows) cudaFreeHost(a->h_hw_rows);
    for (b2p_arena::IoSlot &sl : a->io) {
        if (sl.dev) cudaFree(sl.dev);
        if (sl.dev_jobs) cudaFree(sl.dev_jobs);
        if (sl.snap) cudaEventDestroy(sl.snap);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    if (a->io_stream) cudaStreamDestroy(a->io_stream);
    if (a->up_stream) cudaStreamDestroy(a->up_stream);
    if (a->cap_stream) cudaStreamDestroy(a->cap_stream);
    if (a->ev_asm_done) cudaEventDestroy(a->ev_asm_done);
    if (a->ev_h2d_done) cudaEventDestroy(a->ev_h2d_done);
    delete a;
}

int b2p_arena_num_worlds(const b2p_arena *a) { return a ? a->W : 0; }
int b2p_padded_worlds(const b2p_arena *a) { return a ? a->Wp : 0; }

int b2p_arena_set_stream(b2p_arena *a, void *s)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    cudaStreamSynchronize(a->stream);
    a->stream = (cudaStream_t)s;
    return 0;
}

int b2p_arena_set_max_contacts(b2p_arena *a, int n)
{
    if (!a || n < 0) return fail(B2P_EINVAL, "bad max_contacts");
    if (a->stepped) return fail(B2P_EINVAL, "max_contacts must be set before the first step");
    a->maxc = n;
    a->tmpl_dirty = true;
    return 0;
}

int b2p_arena_enable_rolling_friction(b2p_arena *a, int on)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    if (a->stepped) return fail(B2P_EINVAL, "rolling friction must be selected before the first step");
    a->rolling = on ? 1 : 0;
    a->tmpl_dirty = true;
    return 0;
}

int b2p_world_set_gravity(b2p_arena *a, double x, double y, double z)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->gravity[0] = x, a->gravity[1] = y, a->gravity[2] = z;
    return 0;
}
int b2p_world_set_cfm(b2p_arena *a, double v) { return a ? (a->cfm = v, 0) : fail(B2P_EINVAL, "null arena"); }
int b2p_world_set_erp(b2p_arena *a, double v) { return a ? (a->erp = v, 0) : fail(B2P_EINVAL, "null arena"); }
int b2p_world_set_max_angular_speed(b2p_arena *a, double v)
{
    return a ? (a->max_ang = v, 0) : fail(B2P_EINVAL, "null arena");
}
int b2p_world_set_contact_max_correcting_vel(b2p_arena *a, double v)
{
    return a ? (a->contact_max_vel = v, 0) : fail(B2P_EINVAL, "null arena");
}
int b2p_world_set_contact_surface_layer(b2p_arena *a, double v)
{
    return a ? (a->min_depth = v, 0) : fail(B2P_EINVAL, "null arena");
}
int b2p_world_set_quickstep_num_iterations(b2p_arena *a, int n)
{
    if (!a || n < 0) return fail(B2P_EINVAL, "bad iteration count");
    a->iters = n;
    return 0;
}
int b2p_world_set_quickstep_w(b2p_arena *a, double v) { return a ? (a->sor_w = v, 0) : fail(B2P_EINVAL, "null arena"); }
int b2p_world_set_reorder_method(b2p_arena *a, int m)
{
    if (!a || (m != B2P_REORDER_RANDOMLY && m != B2P_REORDER_NONE)) return fail(B2P_EINVAL, "bad reorder method");
    a->reorder = m;
    return 0;
}
int b2p_world_set_row_order(b2p_arena *a, int order)
{
    if (!a || (order != B2P_ORDER_WORLD_BLOCK && order != B2P_ORDER_ODE_ISLANDS)) return fail(B2P_EINVAL, "bad row order");
    a->order_mode = order;
    return 0;
}
int b2p_world_rand_set_seed(b2p_arena *a, uint32_t seed)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->seed.dev_newer = false;
    mset(a, a->seed, 0, B2P_ALL_WORLDS, seed);
    return 0;
}
int b2p_world_rand_get_seed(b2p_arena *a, int world, uint32_t *seed)
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    if ((rc = mirror_fetch_column(a, a->seed, 0, 1, world))) return rc;
    *seed = a->seed.host[(size_t)world];
    return 0;
}

// ---------------- bodies ----------------
int b2p_body_create(b2p_arena *a)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    if ((int)a->bodies.size() >= B2P_MAX_BODIES) return fail(B2P_ECAPACITY, "too many bodies");
    HBody b;
    b.has_max_ang = a->max_ang < kInf; // inherited from the world at creation (dBodyCreate)
    b.max_angular_speed = a->max_ang;
    a->bodies.push_back(b);
    const int id = (int)a->bodies.size() - 1;
    int rc;
    if ((rc = mirror_resize(a, a->state, (size_t)(id + 1) * B2P_STATE_FIELDS, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->force, (size_t)(id + 1) * 6, 0.f))) return rc;
    mset(a, a->state, (size_t)id * B2P_STATE_FIELDS + 3, B2P_ALL_WORLDS, 1.0f);
    // padded lanes get a valid quaternion too
    for (int w = a->W; w < a->Wp; w++) a->state.host[((size_t)id * B2P_STATE_FIELDS + 3) * a->Wp + w] = 1.0f;
    a->tmpl_dirty = true;
    return id;
}

int b2p_body_set_mass(b2p_arena *a, int body, double mass, const double I[9])
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if (!(mass > 0)) return fail(B2P_EINVAL, "mass must be positive");
    HBody &b = a->bodies[body];
    b.mass = mass;
    memcpy(b.I, I, sizeof(b.I));
    if (!invert33(b.invI, b.I)) return fail(B2P_EINVAL, "inertia tensor is singular");
    a->tmpl_dirty = true;
    return 0;
}

static int set3(b2p_arena *a, Mirror<float> &m, size_t field0, int body, int world, double x, double y, double z,
                bool add)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if ((rc = check_world(a, world, true))) return rc;
    const double v[3] = {x, y, z};
    if (world != B2P_ALL_WORLDS && m.dev_newer) {
        // the device copy is the current one: update this world's column in place (read-modify-write for `add`)
        if (add && (rc = mirror_fetch_column(a, m, field0, 3, world))) return rc;
        for (int i = 0; i < 3; i++) {
            float &x0 = m.host[(field0 + i) * (size_t)a->Wp + world];
            x0 = add ? x0 + (float)v[i] : (float)v[i];
        }
        return mirror_push_column(a, m, field0, 3, world);
    }
    if ((rc = mirror_to_host(a, m))) return rc;
    for (int i = 0; i < 3; i++) {
        float *row = m.host.data() + (field0 + i) * (size_t)a->Wp;
        if (world == B2P_ALL_WORLDS) {
            for (int w = 0; w < a->W; w++) row[w] = add ? row[w] + (float)v[i] : (float)v[i];
        } else {
            row[world] = add ? row[world] + (float)v[i] : (float)v[i];
        }
    }
    m.host_newer = true;
    return 0;
}

int b2p_body_set_position(b2p_arena *a, int body, int world, double x, double y, double z)
{
    int rc = set3(a, a->state, (size_t)body * B2P_STATE_FIELDS + 0, body, world, x, y, z, false);
    if (rc) return rc;
    if (world == B2P_ALL_WORLDS || world == 0) {
        HBody &b = a->bodies[body];
        b.pos[0] = x, b.pos[1] = y, b.pos[2] = z;
    }
    return 0;
}

int b2p_body_set_quaternion(b2p_arena *a, int body, int world, double qw, double qx, double qy, double qz)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if ((rc = check_world(a, world, true))) return rc;
    double q[4] = {qw, qx, qy, qz};
    norm4(q); // dBodySetQuaternion normalises
    if (world == B2P_ALL_WORLDS || world == 0) memcpy(a->bodies[body].q, q, sizeof(q));
    const size_t f0 = (size_t)body * B2P_STATE_FIELDS + 3;
    if (world != B2P_ALL_WORLDS && a->state.dev_newer) {
        for (int i = 0; i < 4; i++) a->state.host[(f0 + i) * (size_t)a->Wp + world] = (float)q[i];
        return mirror_push_column(a, a->state, f0, 4, world);
    }
    if ((rc = mirror_to_host(a, a->state))) return rc;
    for (int i = 0; i < 4; i++) mset(a, a->state, f0 + i, world, (float)q[i]);
    return 0;
}

int b2p_body_set_linear_vel(b2p_arena *a, int body, int world, double x, double y, double z)
{
    return set3(a, a->state, (size_t)body * B2P_STATE_FIELDS + 7, body, world, x, y, z, false);
}
int b2p_body_set_angular_vel(b2p_arena *a, int body, int world, double x, double y, double z)
{
    return set3(a, a->state, (size_t)body * B2P_STATE_FIELDS + 10, body, world, x, y, z, false);
}
int b2p_body_set_force(b2p_arena *a, int body, int world, double x, double y, double z)
{
    return set3(a, a->force, (size_t)body * 6 + 0, body, world, x, y, z, false);
}
int b2p_body_set_torque(b2p_arena *a, int body, int world, double x, double y, double z)
{
    return set3(a, a->force, (size_t)body * 6 + 3, body, world, x, y, z, false);
}
int b2p_body_add_force(b2p_arena *a, int body, int world, double x, double y, double z)
{
    return set3(a, a->force, (size_t)body * 6 + 0, body, world, x, y, z, true);
}
int b2p_body_add_torque(b2p_arena *a, int body, int world, double x, double y, double z)
{
    return set3(a, a->force, (size_t)body * 6 + 3, body, world, x, y, z, true);
}
int b2p_body_add_force_at_pos(b2p_arena *a, int body, int world, double fx, double fy, double fz, double px,
                              double py, double pz)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if ((rc = check_world(a, world, true))) return rc;
    const int w0 = world == B2P_ALL_WORLDS ? 0 : world, w1 = world == B2P_ALL_WORLDS ? a->W : world + 1;
    for (int w = w0; w < w1; w++) {
        double pos[3], q[4];
        if ((rc = body_pose(a, body, w, pos, q))) return rc;
        const double f[3] = {fx, fy, fz}, r[3] = {px - pos[0], py - pos[1], pz - pos[2]};
        double t[3];
        cross(t, r, f);
        if ((rc = set3(a, a->force, (size_t)body * 6 + 0, body, w, f[0], f[1], f[2], true))) return rc;
        if ((rc = set3(a, a->force, (size_t)body * 6 + 3, body, w, t[0], t[1], t[2], true))) return rc;
    }
    return 0;
}
int b2p_body_set_gravity_mode(b2p_arena *a, int body, int on)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    a->bodies[body].nogravity = !on;
    a->tmpl_dirty = true;
    return 0;
}
int b2p_body_set_gyroscopic_mode(b2p_arena *a, int body, int on)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    a->bodies[body].gyroscopic = on ? 1 : 0;
    a->tmpl_dirty = true;
    return 0;
}

static int get_n(b2p_arena *a, Mirror<float> &m, size_t field0, int body, int world, int n, double *out)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    return body_vec(a, m, field0, world, n, out);
}
int b2p_body_get_position(b2p_arena *a, int body, int world, double out[3])
{
    return get_n(a, a->state, (size_t)body * B2P_STATE_FIELDS + 0, body, world, 3, out);
}
int b2p_body_get_quaternion(b2p_arena *a, int body, int world, double out[4])
{
    return get_n(a, a->state, (size_t)body * B2P_STATE_FIELDS + 3, body, world, 4, out);
}
int b2p_body_get_rotation(b2p_arena *a, int body, int world, double out[9])
{
    double q[4];
    int rc = b2p_body_get_quaternion(a, body, world, q);
    if (rc) return rc;
    q_to_R(q, out);
    return 0;
}
int b2p_body_get_linear_vel(b2p_arena *a, int body, int world, double out[3])
{
    return get_n(a, a->state, (size_t)body * B2P_STATE_FIELDS + 7, body, world, 3, out);
}
int b2p_body_get_angular_vel(b2p_arena *a, int body, int world, double out[3])
{
    return get_n(a, a->state, (size_t)body * B2P_STATE_FIELDS + 10, body, world, 3, out);
}
int b2p_body_get_force(b2p_arena *a, int body, int world, double out[3])
{
    return get_n(a, a->force, (size_t)body * 6 + 0, body, world, 3, out);
}
int b2p_body_get_torque(b2p_arena *a, int body, int world, double out[3])
{
    return get_n(a, a->force, (size_t)body * 6 + 3, body, world, 3, out);
}
int b2p_body_get_state(b2p_arena *a, int body, int world, double out[13])
{
    return get_n(a, a->state, (size_t)body * B2P_STATE_FIELDS, body, world, 13, out);
}
int b2p_body_set_state(b2p_arena *a, int body, int world, const double in[13])
{
    int rc;
    if ((rc = b2p_body_set_position(a, body, world, in[0], in[1], in[2]))) return rc;
    if ((rc = b2p_body_set_quaternion(a, body, world, in[3], in[4], in[5], in[6]))) return rc;
    if ((rc = b2p_body_set_linear_vel(a, body, world, in[7], in[8], in[9]))) return rc;
    return b2p_body_set_angular_vel(a, body, world, in[10], in[11], in[12]);
}

int b2p_state_upload(b2p_arena *a, const float *host_state)
{
    if (!a || !host_state) return fail(B2P_EINVAL, "null argument");
    const size_t nf = a->bodies.size() * B2P_STATE_FIELDS;
    int rc = mirror_to_device(a, a->state); // allocates
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(a->state.dev, sizeof(float) * a->Wp, host_state, sizeof(float) * a->W,
                               sizeof(float) * a->W, nf, cudaMemcpyHostToDevice, a->stream));
    a->state.dev_newer = true;
    a->state.host_newer = false;
    return 0;
}

int b2p_state_download(b2p_arena *a, float *host_state)
{
    if (!a || !host_state) return fail(B2P_EINVAL, "null argument");
    const size_t nf = a->bodies.size() * B2P_STATE_FIELDS;
    int rc = mirror_to_device(a, a->state);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(host_state, sizeof(float) * a->W, a->state.dev, sizeof(float) * a->Wp,
                               sizeof(float) * a->W, nf, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

// The joint observations the integrate kernel wrote after the last step: float[nj][4][num_worlds] =
// angle1, rate1, angle2, rate2 per joint (dJointGet*Angle / *Rate for every world at once).
int b2p_joint_obs_download(b2p_arena *a, float *host_obs)
{
    if (!a || !host_obs) return fail(B2P_EINVAL, "null argument");
    if (!a->stepped || !a->jobs.dev) return fail(B2P_EINVAL, "no step has run yet");
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaMemcpy2DAsync(host_obs, sizeof(float) * a->W, a->jobs.dev, sizeof(float) * a->Wp, sizeof(float) * a->W,
                               a->joints.size() * 4, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

int b2p_body_get_state_batch(b2p_arena *a, int body, float *host_out)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if (!host_out) return fail(B2P_EINVAL, "host_out is null");
    if ((rc = mirror_to_device(a, a->state))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(host_out, sizeof(float) * a->W,
                               a->state.dev + (size_t)body * B2P_STATE_FIELDS * a->Wp, sizeof(float) * a->Wp,
                               sizeof(float) * a->W, B2P_STATE_FIELDS, cudaMemcpyDeviceToHost, a->stream));
    return 0;
}

// Pipelined variant: snapshots the body's state on the compute stream (device-to-device, a few microseconds),
// then copies the snapshot to the host on a separate copy stream, so the transfer overlaps the next step.
// Returns a ticket (>= 0) for b2p_io_wait, which blocks until the host buffer is complete.
int b2p_observe_begin(b2p_arena *a, int body, float *host_out, float *host_joint_obs)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if (!host_out) return fail(B2P_EINVAL, "host_out is null");
    CUDA_TRY(cudaSetDevice(a->device));
    if ((rc = mirror_to_device(a, a->state))) return rc;
    const size_t nj = a->joints.size();
    if (host_joint_obs && (!a->stepped || !a->jobs.dev || nj == 0)) return fail(B2P_EINVAL, "no joint observations yet");
    if (!a->io_stream) CUDA_TRY(cudaStreamCreateWithFlags(&a->io_stream, cudaStreamNonBlocking));
    const int t = a->io_next;
    a->io_next ^= 1;
    b2p_arena::IoSlot &sl = a->io[t];
    const size_t elems = (size_t)B2P_STATE_FIELDS * a->Wp, jelems = host_joint_obs ? nj * 4 * a->Wp : 0;
    if (sl.elems != elems) {
        if (sl.dev) CUDA_TRY(cudaFree(sl.dev));
        CUDA_TRY(cudaMalloc(&sl.dev, sizeof(float) * elems));
        sl.elems = elems;
    }
    if (jelems && sl.jobs_elems != jelems) {
        if (sl.dev_jobs) CUDA_TRY(cudaFree(sl.dev_jobs));
        CUDA_TRY(cudaMalloc(&sl.dev_jobs, sizeof(float) * jelems));
        sl.jobs_elems = jelems;
    }
    if (!sl.snap) {
        CUDA_TRY(cudaEventCreateWithFlags(&sl.snap, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    // the slot's previous transfer must have left the staging buffer before it is overwritten
    if (sl.pending) CUDA_TRY(cudaStreamWaitEvent(a->stream, sl.done, 0));
    CUDA_TRY(cudaMemcpyAsync(sl.dev, a->state.dev + (size_t)body * B2P_STATE_FIELDS * a->Wp, sizeof(float) * elems,
                             cudaMemcpyDeviceToDevice, a->stream));
    if (jelems)
        CUDA_TRY(cudaMemcpyAsync(sl.dev_jobs, a->jobs.dev, sizeof(float) * jelems, cudaMemcpyDeviceToDevice, a->stream));
    CUDA_TRY(cudaEventRecord(sl.snap, a->stream));
    CUDA_TRY(cudaStreamWaitEvent(a->io_stream, sl.snap, 0));
    if (a->W == a->Wp) { // no padding between the fields: one contiguous transfer each
        CUDA_TRY(cudaMemcpyAsync(host_out, sl.dev, sizeof(float) * elems, cudaMemcpyDeviceToHost, a->io_stream));
        if (jelems)
            CUDA_TRY(cudaMemcpyAsync(host_joint_obs, sl.dev_jobs, sizeof(float) * jelems, cudaMemcpyDeviceToHost, a->io_stream));
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(host_out, sizeof(float) * a->W, sl.dev, sizeof(float) * a->Wp, sizeof(float) * a->W,
                                   B2P_STATE_FIELDS, cudaMemcpyDeviceToHost, a->io_stream));
        if (jelems)
            CUDA_TRY(cudaMemcpy2DAsync(host_joint_obs, sizeof(float) * a->W, sl.dev_jobs, sizeof(float) * a->Wp,
                                       sizeof(float) * a->W, nj * 4, cudaMemcpyDeviceToHost, a->io_stream));
    }
    CUDA_TRY(cudaEventRecord(sl.done, a->io_stream));
    sl.pending = true;
    return t;
}

int b2p_body_get_state_batch_begin(b2p_arena *a, int body, float *host_out)
{
    return b2p_observe_begin(a, body, host_out, nullptr);
}

int b2p_io_wait(b2p_arena *a, int ticket)
{
    if (!a || ticket < 0 || ticket > 1) return fail(B2P_EINVAL, "bad ticket");
    b2p_arena::IoSlot &sl = a->io[ticket];
    if (!sl.pending) return 0;
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaEventSynchronize(sl.done));
    return 0;
}

void *b2p_device_state(b2p_arena *a) { return a ? a->state.dev : nullptr; }
void *b2p_device_joint_cmd(b2p_arena *a) { return a ? a->jcmd.dev : nullptr; }
void *b2p_device_joint_obs(b2p_arena *a) { return a ? a->jobs.dev : nullptr; }
void *b2p_device_joint_feedback(b2p_arena *a) { return a ? a->jfb.dev : nullptr; }
void *b2p_device_contact_records(b2p_arena *a) { return a ? a->crec.dev : nullptr; }
void *b2p_device_contact_count(b2p_arena *a) { return a ? a->ccount.dev : nullptr; }

// dJointFeedback of every joint and world after the last step: float[nj][13][num_worlds] = f1 t1 f2 t2 lambda
int b2p_joint_feedback_download(b2p_arena *a, float *host_fb)
{
    if (!a || !host_fb) return fail(B2P_EINVAL, "null argument");
    if (!a->stepped || !a->jfb.dev) return fail(B2P_EINVAL, "no step has run yet");
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaMemcpy2DAsync(host_fb, sizeof(float) * a->W, a->jfb.dev, sizeof(float) * a->Wp, sizeof(float) * a->W,
                               a->joints.size() * 13, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

// ---------------- joints ----------------
int b2p_joint_create(b2p_arena *a, int type, int body1, int body2)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    if (type != B2P_JOINT_HINGE && type != B2P_JOINT_HINGE2 && type != B2P_JOINT_SLIDER && type != B2P_JOINT_FIXED &&
        type != B2P_JOINT_BALL && type != B2P_JOINT_UNIVERSAL)
        return fail(B2P_EUNSUPPORTED, "joint type %d is not supported", type);
    if (body1 >= (int)a->bodies.size() || body2 >= (int)a->bodies.size() || body1 < -1 || body2 < -1)
        return fail(B2P_EINVAL, "joint body id out of range");
    if (type == B2P_JOINT_HINGE2 && (body1 < 0 || body2 < 0))
        return fail(B2P_EINVAL, "a hinge2 joint needs two bodies");
    if ((int)a->joints.size() >= B2P_MAX_JOINTS) return fail(B2P_ECAPACITY, "too many joints");
    HJoint j;
    j.type = type;
    if (body1 < 0 && body2 >= 0) { // dJointAttach: swap and flag dJOINT_REVERSE
        j.b[0] = body2;
        j.b[1] = -1;
        j.reverse = 1;
    } else {
        j.b[0] = body1;
        j.b[1] = body2;
    }
    HLimot l;
    l.normal_cfm = a->cfm;
    l.stop_erp = a->erp;
    l.stop_cfm = a->cfm;
    j.limot1 = j.limot2 = l;
    j.erp = a->erp;
    j.cfm = a->cfm;
    j.susp_erp = a->erp;
    j.susp_cfm = a->cfm;
    if (type == B2P_JOINT_HINGE2 || type == B2P_JOINT_UNIVERSAL) {
        j.axis2[0] = 0;
        j.axis2[1] = 1;
    }
    a->joints.push_back(j);
    const int id = (int)a->joints.size() - 1;
    int rc = mirror_resize(a, a->jcmd, (size_t)(id + 1) * 4, 0.f);
    if (rc) return rc;
    a->tmpl_dirty = true;
    return id;
}

int b2p_joint_set_anchor(b2p_arena *a, int ji, double x, double y, double z)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    HJoint &j = a->joints[ji];
    if ((rc = refresh_template_poses(a, j))) return rc;
    if (j.type == B2P_JOINT_HINGE) {
        set_anchors(a, j, x, y, z);
        initial_relative_rotation(a, j);
    } else if (j.type == B2P_JOINT_HINGE2) {
        set_anchors(a, j, x, y, z);
        hinge2_make_v(a, j);
        hinge2_make_w(a, j);
    } else if (j.type == B2P_JOINT_BALL) {
        set_anchors(a, j, x, y, z);
    } else if (j.type == B2P_JOINT_UNIVERSAL) {
        set_anchors(a, j, x, y, z);
        universal_initial_rotations(a, j);
    }
    a->tmpl_dirty = true;
    return 0;
}

int b2p_joint_set_axis1(b2p_arena *a, int ji, double x, double y, double z)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    HJoint &j = a->joints[ji];
    if ((rc = refresh_template_poses(a, j))) return rc;
    if (j.type == B2P_JOINT_HINGE) {
        set_axes(a, j, x, y, z, j.axis1, j.axis2);
        initial_relative_rotation(a, j);
    } else if (j.type == B2P_JOINT_SLIDER) {
        set_axes(a, j, x, y, z, j.axis1, nullptr);
        // computeOffset
        if (j.b[1] >= 0) {
            const HBody &b0 = a->bodies[j.b[0]], &b1 = a->bodies[j.b[1]];
            double c[3] = {b0.pos[0] - b1.pos[0], b0.pos[1] - b1.pos[1], b0.pos[2] - b1.pos[2]}, R[9];
            q_to_R(b1.q, R);
            Rtv(j.offset, R, c);
        } else if (j.b[0] >= 0) {
            memcpy(j.offset, a->bodies[j.b[0]].pos, sizeof(j.offset));
        }
        initial_relative_rotation(a, j);
    } else if (j.type == B2P_JOINT_HINGE2) {
        set_axes(a, j, x, y, z, j.axis1, nullptr);
        hinge2_axis_info(a, j);
        hinge2_make_v(a, j);
        hinge2_make_w(a, j);
    } else if (j.type == B2P_JOINT_UNIVERSAL) { // dJointSetUniversalAxis1
        if (j.reverse) set_axes(a, j, x, y, z, nullptr, j.axis2);
        else set_axes(a, j, x, y, z, j.axis1, nullptr);
        universal_initial_rotations(a, j);
    }
    a->tmpl_dirty = true;
    return 0;
}

int b2p_joint_set_axis2(b2p_arena *a, int ji, double x, double y, double z)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    HJoint &j = a->joints[ji];
    if ((rc = refresh_template_poses(a, j))) return rc;
    if (j.type == B2P_JOINT_UNIVERSAL) { // dJointSetUniversalAxis2
        if (j.reverse) set_axes(a, j, x, y, z, j.axis1, nullptr);
        else set_axes(a, j, x, y, z, nullptr, j.axis2);
        universal_initial_rotations(a, j);
        a->tmpl_dirty = true;
        return 0;
    }
    if (j.type == B2P_JOINT_HINGE2) {
        set_axes(a, j, x, y, z, nullptr, j.axis2);
        hinge2_axis_info(a, j);
        hinge2_make_v(a, j);
        hinge2_make_w(a, j);
        a->tmpl_dirty = true;
    }
    return 0;
}

int b2p_joint_set_fixed(b2p_arena *a, int ji)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    HJoint &j = a->joints[ji];
    if (j.b[0] < 0) return 0;
    if ((rc = refresh_template_poses(a, j))) return rc;
    const HBody &b0 = a->bodies[j.b[0]];
    if (j.b[1] >= 0) {
        const HBody &b1 = a->bodies[j.b[1]];
        double ofs[3] = {b0.pos[0] - b1.pos[0], b0.pos[1] - b1.pos[1], b0.pos[2] - b1.pos[2]}, R[9];
        q_to_R(b0.q, R);
        Rtv(j.offset, R, ofs);
    } else {
        memcpy(j.offset, b0.pos, sizeof(j.offset));
    }
    initial_relative_rotation(a, j);
    a->tmpl_dirty = true;
    return 0;
}

int b2p_joint_set_param(b2p_arena *a, int ji, int world, int param, double value)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, true))) return rc;
    HJoint &j = a->joints[ji];
    if (j.type == B2P_JOINT_FIXED || j.type == B2P_JOINT_BALL) { // dJointSetFixedParam / dJointSetBallParam
        if (param == B2P_PARAM_CFM) j.cfm = value;
        else if (param == B2P_PARAM_ERP) j.erp = value;
        a->tmpl_dirty = true;
        return 0;
    }
    if (param == B2P_PARAM_SUSPENSION_ERP || param == B2P_PARAM_SUSPENSION_CFM) {
        (param == B2P_PARAM_SUSPENSION_ERP ? j.susp_erp : j.susp_cfm) = value;
        a->tmpl_dirty = true;
        return 0;
    }
    int num;
    HLimot *l = limot_for(j, param, &num);
    const int axis = (l == &j.limot2) ? 1 : 0;
    if (num == B2P_PARAM_VEL || num == B2P_PARAM_FMAX) {
        if (num == B2P_PARAM_FMAX && value < 0) return 0; // ODE ignores negative fmax
        if (world == B2P_ALL_WORLDS || world == 0) (num == B2P_PARAM_VEL ? l->vel : l->fmax) = value;
        const size_t field = (size_t)ji * 4 + axis * 2 + (num == B2P_PARAM_FMAX ? 1 : 0);
        if (world != B2P_ALL_WORLDS && a->jcmd.dev_newer) { // device copy current: update this world's value in place
            if ((rc = join_pending_upload(a))) return rc;
            a->jcmd.host[field * (size_t)a->Wp + world] = (float)value;
            return mirror_push_column(a, a->jcmd, field, 1, world);
        }
        if ((rc = mirror_to_host(a, a->jcmd))) return rc;
        mset(a, a->jcmd, field, world, (float)value);
        return 0;
    }
    if (world != B2P_ALL_WORLDS && a->W > 1)
        return fail(B2P_EUNSUPPORTED, "only VEL and FMAX can differ between worlds; use B2P_ALL_WORLDS");
    switch (num) {
    case B2P_PARAM_LO_STOP: l->lostop = value; break;
    case B2P_PARAM_HI_STOP: l->histop = value; break;
    case B2P_PARAM_FUDGE_FACTOR:
        if (value >= 0 && value <= 1) l->fudge = value;
        break;
    case B2P_PARAM_BOUNCE: l->bounce = value; break;
    case B2P_PARAM_CFM: l->normal_cfm = value; break;
    case B2P_PARAM_STOP_ERP: l->stop_erp = value; break;
    case B2P_PARAM_STOP_CFM: l->stop_cfm = value; break;
    default: break;
    }
    a->tmpl_dirty = true;
    return 0;
}

int b2p_joint_get_param(b2p_arena *a, int ji, int world, int param, double *value)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if (world == B2P_ALL_WORLDS) world = 0;
    if ((rc = check_world(a, world, false))) return rc;
    HJoint &j = a->joints[ji];
    *value = 0;
    if (j.type == B2P_JOINT_FIXED || j.type == B2P_JOINT_BALL) {
        if (param == B2P_PARAM_CFM) *value = j.cfm;
        if (param == B2P_PARAM_ERP) *value = j.erp;
        return 0;
    }
    if (param == B2P_PARAM_SUSPENSION_ERP) return *value = j.susp_erp, 0;
    if (param == B2P_PARAM_SUSPENSION_CFM) return *value = j.susp_cfm, 0;
    int num;
    HLimot *l = limot_for(j, param, &num);
    const int axis = (l == &j.limot2) ? 1 : 0;
    switch (num) {
    case B2P_PARAM_VEL:
    case B2P_PARAM_FMAX:
    {
        const size_t field = (size_t)ji * 4 + axis * 2 + (num == B2P_PARAM_FMAX ? 1 : 0);
        if ((rc = mirror_fetch_column(a, a->jcmd, field, 1, world))) return rc;
        *value = a->jcmd.host[field * a->Wp + world];
        break;
    }
    case B2P_PARAM_LO_STOP: *value = l->lostop; break;
    case B2P_PARAM_HI_STOP: *value = l->histop; break;
    case B2P_PARAM_FUDGE_FACTOR: *value = l->fudge; break;
    case B2P_PARAM_BOUNCE: *value = l->bounce; break;
    case B2P_PARAM_CFM: *value = l->normal_cfm; break;
    case B2P_PARAM_STOP_ERP: *value = l->stop_erp; break;
    case B2P_PARAM_STOP_CFM: *value = l->stop_cfm; break;
    default: break;
    }
    return 0;
}

int b2p_joint_set_param_batch(b2p_arena *a, int ji, int param, const float *values)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if (!values) return fail(B2P_EINVAL, "values is null");
    const int num = param & 0xff, axis = ((param & 0xff00) == B2P_PARAM_GROUP2) ? 1 : 0;
    if (num != B2P_PARAM_VEL && num != B2P_PARAM_FMAX)
        return fail(B2P_EUNSUPPORTED, "only VEL and FMAX can be set per world");
    if ((rc = join_pending_upload(a))) return rc;
    if ((rc = mirror_to_device(a, a->jcmd))) return rc; // device copy current
    const size_t field = (size_t)ji * 4 + axis * 2 + (num == B2P_PARAM_FMAX ? 1 : 0);
    CUDA_TRY(cudaMemcpyAsync(a->jcmd.dev + field * a->Wp, values, sizeof(float) * a->W, cudaMemcpyHostToDevice,
                             a->stream));
    a->jcmd.dev_newer = true;
    return 0;
}

// One transfer for a whole command vector: values[num_joints][num_worlds] -> the same per-world parameter
// (VEL or FMAX, optionally | GROUP2) of joints 0 .. num_joints-1.
int b2p_joints_set_param_batch(b2p_arena *a, int param, const float *values, int num_joints)
{
    if (!a || !values) return fail(B2P_EINVAL, "null argument");
    if (num_joints <= 0 || num_joints > (int)a->joints.size()) return fail(B2P_EINVAL, "bad joint count");
    const int num = param & 0xff, axis = ((param & 0xff00) == B2P_PARAM_GROUP2) ? 1 : 0;
    if (num != B2P_PARAM_VEL && num != B2P_PARAM_FMAX)
        return fail(B2P_EUNSUPPORTED, "only VEL and FMAX can be set per world");
    CUDA_TRY(cudaSetDevice(a->device));
    // If the device copy is current and a step has been enqueued, the transfer goes over the upload stream behind
    // that step's assemble kernel (the only reader of jcmd), i.e. it overlaps the step's SOR-LCP kernel.
    a->upload_mode = true;
    const bool overlap = !a->jcmd.host_newer && a->jcmd.dev && a->asm_recorded;
    int rc = mirror_to_device(a, a->jcmd);
    if (rc) return rc;
    const size_t field = (size_t)axis * 2 + (num == B2P_PARAM_FMAX ? 1 : 0);
    cudaStream_t st = a->stream;
    if (overlap) {
        if ((rc = join_pending_upload(a))) return rc; // uploads stay in order among themselves
        if (!a->ev_h2d_done) CUDA_TRY(cudaEventCreateWithFlags(&a->ev_h2d_done, cudaEventDisableTiming));
        // (host -> device and device -> host transfers run concurrently on the link: the upload gets its own
        // stream instead of waiting behind the observation of the previous step on the read-back stream)
        if (!a->up_stream) CUDA_TRY(cudaStreamCreateWithFlags(&a->up_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamWaitEvent(a->up_stream, a->ev_asm_done, 0));
        st = a->up_stream;
    }
    CUDA_TRY(cudaMemcpy2DAsync(a->jcmd.dev + field * a->Wp, sizeof(float) * 4 * a->Wp, values, sizeof(float) * a->W,
                               sizeof(float) * a->W, (size_t)num_joints, cudaMemcpyHostToDevice, st));
    if (overlap) {
        CUDA_TRY(cudaEventRecord(a->ev_h2d_done, a->up_stream));
        a->h2d_pending = true;
    }
    a->jcmd.dev_newer = true;
    return 0;
}

static int joint_poses(b2p_arena *a, const HJoint &j, int world, double *p0, double *q0, double *p1, double *q1)
{
    int rc;
    if (j.b[0] >= 0 && (rc = body_pose(a, j.b[0], world, p0, q0))) return rc;
    if (j.b[1] >= 0 && (rc = body_pose(a, j.b[1], world, p1, q1))) return rc;
    return 0;
}

int b2p_joint_get_anchor(b2p_arena *a, int ji, int world, double out[3])
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    out[0] = out[1] = out[2] = 0;
    if (j.b[0] < 0) return 0;
    double p0[3], q0[4], p1[3], q1[4], R[9];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q0, R);
    Rv(out, R, j.anchor1);
    for (int i = 0; i < 3; i++) out[i] += p0[i];
    return 0;
}

int b2p_joint_get_anchor2(b2p_arena *a, int ji, int world, double out[3])
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    memcpy(out, j.anchor2, 3 * sizeof(double));
    if (j.b[1] < 0) return 0;
    double p0[3], q0[4], p1[3], q1[4], R[9];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q1, R);
    Rv(out, R, j.anchor2);
    for (int i = 0; i < 3; i++) out[i] += p1[i];
    return 0;
}

int b2p_joint_get_axis1(b2p_arena *a, int ji, int world, double out[3])
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    out[0] = out[1] = out[2] = 0;
    if (j.b[0] < 0) return 0;
    double p0[3], q0[4], p1[3], q1[4], R[9];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q0, R);
    Rv(out, R, j.axis1);
    if (j.type == B2P_JOINT_SLIDER && j.reverse)
        for (int i = 0; i < 3; i++) out[i] = -out[i];
    return 0;
}

int b2p_joint_get_axis2(b2p_arena *a, int ji, int world, double out[3])
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    out[0] = out[1] = out[2] = 0;
    if (j.b[1] < 0) return 0;
    double p0[3], q0[4], p1[3], q1[4], R[9];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q1, R);
    Rv(out, R, j.axis2);
    return 0;
}

int b2p_joint_get_angle1(b2p_arena *a, int ji, int world, double *out)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    *out = 0;
    if (j.b[0] < 0) return 0;
    double p0[3], q0[4], p1[3] = {0, 0, 0}, q1[4] = {1, 0, 0, 0}, R0[9], R1[9];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q0, R0);
    q_to_R(q1, R1);
    if (j.type == B2P_JOINT_HINGE) {
        double ang = hinge_angle_host(j, q0, q1);
        *out = j.reverse ? -ang : ang;
    } else if (j.type == B2P_JOINT_SLIDER) {
        double ax1[3], q[3];
        Rv(ax1, R0, j.axis1);
        if (j.b[1] >= 0) {
            Rv(q, R1, j.offset);
            for (int i = 0; i < 3; i++) q[i] = p0[i] - q[i] - p1[i];
        } else {
            for (int i = 0; i < 3; i++) q[i] = p0[i] - j.offset[i];
            if (j.reverse)
                for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
        }
        *out = dot(ax1, q);
    } else if (j.type == B2P_JOINT_HINGE2) {
        double p[3], q[3];
        Rv(p, R1, j.axis2);
        Rtv(q, R0, p);
        *out = -std::atan2(dot(j.v2, q), dot(j.v1, q));
    } else if (j.type == B2P_JOINT_UNIVERSAL) { // dJointGetUniversalAngle1
        double ax1[3], ax2[3], a1, a2;
        universal_axes_world(j, q0, q1, ax1, ax2);
        b2p_uni::angles<double>(ax1, ax2, q0, q1, j.b[1] >= 0, j.axis1, j.axis2, j.qrel, j.qrel2, &a1, &a2);
        *out = j.reverse ? -a2 : a1;
    }
    return 0;
}

int b2p_joint_get_angle1_rate(b2p_arena *a, int ji, int world, double *out)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    *out = 0;
    if (j.b[0] < 0 || j.type == B2P_JOINT_FIXED || j.type == B2P_JOINT_BALL) return 0;
    double p0[3], q0[4], p1[3], q1[4] = {1, 0, 0, 0}, R0[9], axis[3], lv0[3], av0[3], lv1[3] = {0, 0, 0}, av1[3] = {0, 0, 0};
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q0, R0);
    Rv(axis, R0, j.axis1);
    b2p_body_get_linear_vel(a, j.b[0], world, lv0);
    b2p_body_get_angular_vel(a, j.b[0], world, av0);
    if (j.b[1] >= 0) {
        b2p_body_get_linear_vel(a, j.b[1], world, lv1);
        b2p_body_get_angular_vel(a, j.b[1], world, av1);
    }
    if (j.type == B2P_JOINT_UNIVERSAL) { // dJointGetUniversalAngle1Rate
        double ax1[3], ax2[3];
        universal_axes_world(j, q0, q1, ax1, ax2);
        const double *ax = j.reverse ? ax2 : ax1;
        *out = dot(ax, av0) - dot(ax, av1);
    } else if (j.type == B2P_JOINT_HINGE || j.type == B2P_JOINT_HINGE2) {
        double rate = dot(axis, av0);
        if (j.b[1] >= 0) rate -= dot(axis, av1);
        if (j.type == B2P_JOINT_HINGE && j.reverse) rate = -rate;
        *out = rate;
    } else if (j.type == B2P_JOINT_SLIDER) {
        if (j.b[1] >= 0) {
            *out = dot(axis, lv0) - dot(axis, lv1);
        } else {
            double rate = dot(axis, lv0);
            *out = j.reverse ? -rate : rate;
        }
    }
    return 0;
}

int b2p_joint_get_angle2(b2p_arena *a, int ji, int world, double *out)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    *out = 0;
    if (j.type == B2P_JOINT_UNIVERSAL && j.b[0] >= 0) { // dJointGetUniversalAngle2
        double p0[3], q0[4], p1[3], q1[4] = {1, 0, 0, 0}, ax1[3], ax2[3], a1, a2;
        if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
        universal_axes_world(j, q0, q1, ax1, ax2);
        b2p_uni::angles<double>(ax1, ax2, q0, q1, j.b[1] >= 0, j.axis1, j.axis2, j.qrel, j.qrel2, &a1, &a2);
        *out = j.reverse ? -a1 : a2;
        return 0;
    }
    if (j.type != B2P_JOINT_HINGE2) return 0;
    double p0[3], q0[4], p1[3], q1[4], R0[9], R1[9], p[3], q[3];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q0, R0);
    q_to_R(q1, R1);
    Rv(p, R0, j.axis1);
    Rtv(q, R1, p);
    *out = -std::atan2(dot(j.w2, q), dot(j.w1, q));
    return 0;
}

int b2p_joint_get_angle2_rate(b2p_arena *a, int ji, int world, double *out)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    const HJoint &j = a->joints[ji];
    *out = 0;
    if (j.type == B2P_JOINT_UNIVERSAL && j.b[0] >= 0) { // dJointGetUniversalAngle2Rate
        double p0[3], q0[4], p1[3], q1[4] = {1, 0, 0, 0}, ax1[3], ax2[3], av0[3], av1[3] = {0, 0, 0};
        if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
        universal_axes_world(j, q0, q1, ax1, ax2);
        b2p_body_get_angular_vel(a, j.b[0], world, av0);
        if (j.b[1] >= 0) b2p_body_get_angular_vel(a, j.b[1], world, av1);
        const double *ax = j.reverse ? ax1 : ax2;
        *out = dot(ax, av0) - dot(ax, av1);
        return 0;
    }
    if (j.type != B2P_JOINT_HINGE2) return 0;
    double p0[3], q0[4], p1[3], q1[4], R1[9], axis[3], av0[3], av1[3];
    if ((rc = joint_poses(a, j, world, p0, q0, p1, q1))) return rc;
    q_to_R(q1, R1);
    Rv(axis, R1, j.axis2);
    b2p_body_get_angular_vel(a, j.b[0], world, av0);
    b2p_body_get_angular_vel(a, j.b[1], world, av1);
    *out = dot(axis, av0) - dot(axis, av1);
    return 0;
}

int b2p_joint_add_torque(b2p_arena *a, int ji, int world, double t)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, true))) return rc;
    const HJoint &j = a->joints[ji];
    if (j.b[0] < 0) return 0;
    if (j.type != B2P_JOINT_HINGE && j.type != B2P_JOINT_SLIDER)
        return fail(B2P_EUNSUPPORTED, "add_torque is defined for hinge and slider joints");
    const int w0 = world == B2P_ALL_WORLDS ? 0 : world, w1 = world == B2P_ALL_WORLDS ? a->W : world + 1;
    for (int w = w0; w < w1; w++) {
        double p0[3], q0[4], p1[3], q1[4], R0[9], axis[3];
        if ((rc = joint_poses(a, j, w, p0, q0, p1, q1))) return rc;
        q_to_R(q0, R0);
        Rv(axis, R0, j.axis1);
        const double torque = j.reverse ? -t : t;
        for (int i = 0; i < 3; i++) axis[i] *= torque;
        if (j.type == B2P_JOINT_HINGE) {
            b2p_body_add_torque(a, j.b[0], w, axis[0], axis[1], axis[2]);
            if (j.b[1] >= 0) b2p_body_add_torque(a, j.b[1], w, -axis[0], -axis[1], -axis[2]);
        } else {
            if (j.b[1] >= 0) {
                double c[3], ltd[3];
                for (int i = 0; i < 3; i++) c[i] = 0.5 * (p1[i] - p0[i]);
                cross(ltd, c, axis);
                b2p_body_add_torque(a, j.b[0], w, ltd[0], ltd[1], ltd[2]);
                b2p_body_add_torque(a, j.b[1], w, ltd[0], ltd[1], ltd[2]);
                b2p_body_add_force(a, j.b[1], w, -axis[0], -axis[1], -axis[2]);
            }
            b2p_body_add_force(a, j.b[0], w, axis[0], axis[1], axis[2]);
        }
    }
    return 0;
}

int b2p_joint_get_feedback(b2p_arena *a, int ji, int world, double out[13])
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    for (int i = 0; i < 13; i++) out[i] = 0;
    if (a->jfb.fields < (size_t)(ji + 1) * 13) return 0; // never stepped
    return body_vec(a, a->jfb, (size_t)ji * 13, world, 13, out);
}

int b2p_joint_get_num_rows(b2p_arena *a, int ji, int world, int *rows)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    *rows = 0;
    if (a->jrows.fields <= (size_t)ji) return 0;
    if ((rc = mirror_fetch_column(a, a->jrows, (size_t)ji, 1, world))) return rc;
    *rows = a->jrows.host[(size_t)ji * a->Wp + world] & 7; // bits 3, 4: which limit / motor rows exist (device use)
    return 0;
}

int b2p_are_connected_excluding_contacts(b2p_arena *a, int b1, int b2)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    if (b1 < 0 || b2 < 0) return 0;
    for (const HJoint &j : a->joints)
        if ((j.b[0] == b1 && j.b[1] == b2) || (j.b[0] == b2 && j.b[1] == b1)) return 1;
    return 0;
}

// ---------------- contacts ----------------
static int ensure_contact_mirrors(b2p_arena *a)
{
    int rc;
    if ((rc = mirror_resize(a, a->crec, (size_t)a->maxc * B2P_CONTACT_FIELDS, 0.f))) return rc;
    if ((rc = mirror_resize(a, a->ccount, 1, 0))) return rc;
    if ((rc = mirror_to_host(a, a->crec))) return rc;
    return mirror_to_host(a, a->ccount);
}

int b2p_contacts_clear(b2p_arena *a, int world)
{
    int rc = check_world(a, world, true);
    if (rc) return rc;
    if (world == B2P_ALL_WORLDS) {
        // "reset the contact counter": no record stays valid, so a device-side stream is simply dropped, and when
        // the counters live on the device they are zeroed there (O(1) on the host)
        if ((rc = mirror_resize(a, a->crec, (size_t)a->maxc * B2P_CONTACT_FIELDS, 0.f))) return rc;
        if ((rc = mirror_resize(a, a->ccount, 1, 0))) return rc;
        a->crec.dev_newer = false;
        if (a->ccount.dev && a->ccount.dev_elems == a->ccount.host.size() && !a->ccount.host_newer) {
            CUDA_TRY(cudaMemsetAsync(a->ccount.dev, 0, sizeof(int) * a->ccount.dev_elems, a->stream));
            a->ccount.dev_newer = true;
            return 0;
        }
        a->ccount.dev_newer = false;
        mset(a, a->ccount, 0, world, 0);
        return 0;
    }
    if ((rc = ensure_contact_mirrors(a))) return rc;
    mset(a, a->ccount, 0, world, 0);
    return 0;
}

int b2p_contact_add(b2p_arena *a, int world, int body1, int body2, const b2p_contact_desc *c)
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    if (!c) return fail(B2P_EINVAL, "contact is null");
    const int nb = (int)a->bodies.size();
    if (body1 >= nb || body2 >= nb || (body1 < 0 && body2 < 0)) return fail(B2P_EINVAL, "bad contact bodies");
    if (body1 >= 0 && body2 >= 0 && b2p_are_connected_excluding_contacts(a, body1, body2))
        return fail(B2P_EREJECTED, "bodies %d and %d are connected by a joint", body1, body2);
    if ((c->mode & B2P_CONTACT_ROLLING) && !a->rolling)
        return fail(B2P_EUNSUPPORTED, "rolling friction needs b2p_arena_enable_rolling_friction(1)");
    if ((rc = ensure_contact_mirrors(a))) return rc;
    int &n = a->ccount.host[(size_t)world];
    if (n >= a->maxc) return fail(B2P_ECAPACITY, "contact capacity (%d) of world %d exhausted", a->maxc, world);
    double sign = 1.0;
    int b0 = body1, b1 = body2;
    if (b0 < 0) { // dJointAttach swap + dJOINT_REVERSE: the normal is negated in getInfo2
        b0 = body2;
        b1 = -1;
        sign = -1.0;
    }
    float rec[B2P_CONTACT_FIELDS];
    const int ids = (b0 & 0xff) | ((b1 < 0 ? 0xff : b1) << 8);
    memcpy(&rec[CR_IDS], &ids, 4);
    for (int i = 0; i < 3; i++) {
        rec[CR_PX + i] = (float)c->pos[i];
        rec[CR_NX + i] = (float)(sign * c->normal[i]);
    }
    rec[CR_DEPTH] = (float)c->depth;
    rec[CR_MU] = (float)c->mu;
    rec[CR_MU2] = (float)c->mu2;
    rec[CR_SOFT_ERP] = (float)c->soft_erp;
    rec[CR_SOFT_CFM] = (float)c->soft_cfm;
    rec[CR_RHO] = (float)c->rho;
    rec[CR_RHO2] = (float)c->rho2;
    rec[CR_RHON] = (float)c->rhoN;
    memcpy(&rec[CR_MODE], &c->mode, 4);
    for (int f = 0; f < B2P_CONTACT_FIELDS; f++)
        a->crec.host[((size_t)n * B2P_CONTACT_FIELDS + f) * a->Wp + world] = rec[f];
    a->crec.host_newer = true;
    a->ccount.host_newer = true;
    return n++;
}

int b2p_contact_count(b2p_arena *a, int world, int *count)
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    *count = 0;
    if (a->ccount.fields == 0) return 0;
    if ((rc = mirror_fetch_column(a, a->ccount, 0, 1, world))) return rc;
    *count = a->ccount.host[(size_t)world];
    return 0;
}

int b2p_contact_get_feedback(b2p_arena *a, int world, int contact, double f1[3])
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    f1[0] = f1[1] = f1[2] = 0;
    if (contact < 0 || contact >= a->maxc) return fail(B2P_EINVAL, "contact id out of range");
    if (a->cfb.fields == 0) return 0;
    return body_vec(a, a->cfb, (size_t)contact * 3, world, 3, f1);
}

int b2p_contact_get(b2p_arena *a, int world, int contact, int *body1, int *body2, b2p_contact_desc *c)
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    if (contact < 0 || contact >= a->maxc || a->crec.fields == 0) return fail(B2P_EINVAL, "contact id out of range");
    if ((rc = mirror_fetch_column(a, a->crec, (size_t)contact * B2P_CONTACT_FIELDS, B2P_CONTACT_FIELDS, world))) return rc;
    float rec[B2P_CONTACT_FIELDS];
    for (int f = 0; f < B2P_CONTACT_FIELDS; f++)
        rec[f] = a->crec.host[((size_t)contact * B2P_CONTACT_FIELDS + f) * a->Wp + world];
    int ids, mode;
    memcpy(&ids, &rec[CR_IDS], 4);
    memcpy(&mode, &rec[CR_MODE], 4);
    if (body1) *body1 = ids & 0xff;
    if (body2) *body2 = ((ids >> 8) & 0xff) == 0xff ? -1 : ((ids >> 8) & 0xff);
    if (c) {
        for (int i = 0; i < 3; i++) {
            c->pos[i] = rec[CR_PX + i];
            c->normal[i] = rec[CR_NX + i];
        }
        c->depth = rec[CR_DEPTH];
        c->mode = mode;
        c->mu = rec[CR_MU];
        c->mu2 = rec[CR_MU2];
        c->rho = rec[CR_RHO];
        c->rho2 = rec[CR_RHO2];
        c->rhoN = rec[CR_RHON];
        c->soft_erp = rec[CR_SOFT_ERP];
        c->soft_cfm = rec[CR_SOFT_CFM];
    }
    return 0;
}

// ---------------- geoms / collision ----------------
int b2p_geom_create(b2p_arena *a, int cls, int body, const double dims[4], const double lpos[3],
                    const double lquat[4], const b2p_surface *s)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    if (cls < B2P_GEOM_SPHERE || cls > B2P_GEOM_PLANE) return fail(B2P_EINVAL, "bad geom class %d", cls);
    if (body >= (int)a->bodies.size() || body < -1) return fail(B2P_EINVAL, "bad geom body");
    if (cls == B2P_GEOM_PLANE && body >= 0) return fail(B2P_EINVAL, "planes are static");
    if (cls != B2P_GEOM_PLANE && body < 0)
        return fail(B2P_EUNSUPPORTED, "static geoms are planes and the heightfield; attach other shapes to a body");
    if ((int)a->geoms.size() >= B2P_MAX_GEOMS) return fail(B2P_ECAPACITY, "too many geoms");
    HGeom g;
    g.cls = cls;
    g.body = body;
    if (dims) memcpy(g.dims, dims, sizeof(g.dims));
    if (lpos) memcpy(g.lpos, lpos, sizeof(g.lpos));
    if (lquat) {
        memcpy(g.lquat, lquat, sizeof(g.lquat));
        norm4(g.lquat);
    }
    if (s) g.s = *s;
    a->geoms.push_back(g);
    a->tmpl_dirty = true;
    return (int)a->geoms.size() - 1;
}

int b2p_geom_create_heightfield(b2p_arena *a, const float *heights, int nx, int ny, double dx, double x0, double y0,
                                const b2p_surface *s)
{
    if (!a || !heights || nx < 2 || ny < 2 || !(dx > 0)) return fail(B2P_EINVAL, "bad heightfield");
    if (a->has_hf) return fail(B2P_ECAPACITY, "one heightfield per arena");
    if ((int)a->geoms.size() >= B2P_MAX_GEOMS) return fail(B2P_ECAPACITY, "too many geoms");
    a->hf_heights.assign(heights, heights + (size_t)nx * ny);
    a->hf_nx = nx;
    a->hf_ny = ny;
    a->hf_dx = dx;
    a->hf_x0 = x0;
    a->hf_y0 = y0;
    a->has_hf = true;
    HGeom g;
    g.cls = B2P_GEOM_HEIGHTFIELD;
    g.body = -1;
    if (s) g.s = *s;
    a->geoms.push_back(g);
    a->tmpl_dirty = true;
    return (int)a->geoms.size() - 1;
}

// ---------------- collision + step: parameter blocks, launches, CUDA graph ----------------
namespace {
constexpr size_t kSmemMax = 227 * 1024;

// Sweep positions per world the SOR kernel keeps in shared memory behind the `kr` positions it keeps in
// registers (one CTA = 16 worlds, 1 KB per position).  Enough for the longest world seen so far (the assemble
// kernel maintains that high-water mark; before the first step the scene's bound `maxrows` is used), capped by
// what one CTA can hold; positions beyond are read from L2 in place, so a stale mark costs time, not
// correctness.  b2p_arena_set_resident_rows overrides the choice.
int pick_rows_resident(const b2p_arena *a, int nb, int maxrows, int kr)
{
    const size_t fixed = b2p_sor_smem_bytes(nb, maxrows, 0), per_row = b2p_sor_smem_bytes(nb, maxrows, 1) - fixed;
    const int cap = kSmemMax - 1024 > fixed ? (int)((kSmemMax - 1024 - fixed) / per_row) : 0;
    const int rest = std::max(maxrows - kr, 0);
    if (a->rows_resident > 0) return std::min(std::min(a->rows_resident, rest), cap);
    const int seen = a->h_hw_rows ? *(volatile int *)a->h_hw_rows : 0;
    const int want = seen > 0 ? std::min(std::max(seen - kr, 0), rest) : rest;
    return std::min(want, cap);
}

// broadphase + narrowphase launch parameters (buffers are created on first use)
int prepare_collide(b2p_arena *a, CollideParams &P, int &sampled)
{
    int rc = finalize(a);
    if (rc) return rc;
    const int ng = (int)a->geoms.size();
    const int maxpairs = ng * (ng - 1) / 2;
    if (a->want_pairs && (maxpairs > a->maxpairs || !a->d_npairs)) {
        if (a->d_pairs) cudaFree(a->d_pairs);
        if (a->d_npairs) cudaFree(a->d_npairs);
        a->d_pairs = a->d_npairs = nullptr;
        CUDA_TRY(cudaMalloc(&a->d_pairs, sizeof(int) * (size_t)std::max(maxpairs, 1) * a->Wp));
        CUDA_TRY(cudaMalloc(&a->d_npairs, sizeof(int) * (size_t)a->Wp));
        a->maxpairs = maxpairs;
    }
    if (!a->d_dropped) {
        CUDA_TRY(cudaMalloc(&a->d_dropped, sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(a->d_dropped, 0, sizeof(int), a->stream));
    }
    memset(&P, 0, sizeof(P));
    P.T = a->dT;
    P.W = a->W;
    P.Wp = a->Wp;
    P.state = a->state.dev;
    P.crec = a->crec.dev;
    P.ccount = a->ccount.dev;
    P.pairs = a->want_pairs ? a->d_pairs : nullptr;
    P.npairs = a->want_pairs ? a->d_npairs : nullptr;
    P.maxpairs = a->want_pairs ? a->maxpairs : 0;
    P.dropped = a->d_dropped;
    if ((size_t)ng * 6 * 128 * sizeof(float) > 200 * 1024) return fail(B2P_ECAPACITY, "too many geoms for the collide kernel");
    // the sample-point colliders (cylinders; capsule against box / heightfield) live in a separate kernel variant
    bool has_cyl = false, has_caps = false, has_box = false;
    for (const HGeom &g : a->geoms) {
        has_cyl |= g.cls == B2P_GEOM_CYLINDER;
        has_caps |= g.cls == B2P_GEOM_CAPSULE;
        has_box |= g.cls == B2P_GEOM_BOX;
    }
    sampled = has_cyl || (has_caps && (has_box || a->has_hf));
    return 0;
}

void after_collide(b2p_arena *a)
{
    a->launches++;
    a->crec.dev_newer = true;
    a->ccount.dev_newer = true;
    a->crec.host_newer = false;
    a->ccount.host_newer = false;
}

// the step's launch parameters and its choice of SOR-LCP kernel
int prepare_step(b2p_arena *a, double step_size, StepParams &P, bool &use_tmem)
{
    int rc = finalize(a);
    if (rc) return rc;
    memset(&P, 0, sizeof(P));
    P.T = a->dT;
    P.W = a->W;
    P.Wp = a->Wp;
    P.h = (float)step_size;
    for (int i = 0; i < 3; i++) P.gravity[i] = (float)a->gravity[i];
    P.erp = (float)a->erp;
    P.cfm = (float)a->cfm;
    P.contact_max_vel = (float)a->contact_max_vel;
    P.contact_min_depth = (float)a->min_depth;
    P.sor_w = (float)a->sor_w;
    P.iters = a->iters;
    P.reorder = a->reorder;
    P.state = a->state.dev;
    P.force = a->force.dev;
    P.jcmd = a->jcmd.dev;
    P.jobs = a->jobs.dev;
    P.jfb = a->jfb.dev;
    P.crec = a->crec.dev;
    P.cfb = a->cfb.dev;
    P.lambda = a->d_lambda;
    P.rowS = a->d_rowS;
    P.invIw = a->d_invIw;
    P.fcacc = a->d_fcacc;
    P.rows = a->d_rows;
    P.cmap = a->d_cmap;
    P.isl = a->d_isl;
    P.ord0 = a->d_ord0;
    P.order_mode = a->order_mode;
    {
        static const int env_tune = getenv("B2P_TUNE") ? atoi(getenv("B2P_TUNE")) : 0;
        P.tune = a->tune >= 0 ? a->tune : env_tune;
    }
    P.jrows = a->jrows.dev;
    P.ccount = a->ccount.dev;
    P.nrows = a->nrows.dev;
    P.seed = a->seed.dev;
    P.bodyx = (a->post_flags & B2P_POST_FRAME_UPDATE) ? a->bodyx.dev : nullptr;
    P.jupd = (a->post_flags & B2P_POST_JOINT_UPDATE) ? a->jupd.dev : nullptr;
    const int nb = (int)a->bodies.size();
    const int maxrows = (int)a->joints.size() * B2P_ROWS_PER_JOINT + a->maxc * rpc_of(a);
    P.hw_rows = a->d_hw_rows;
    // SOR kernel choice.  The Tensor-Memory kernel (one thread per world) wins on every workload measured
    // (configs B, C, D), also when part of a world's rows must be read from L2 in place, so it is used whenever
    // its per-world arrays leave room for an overflow pool in shared memory: with 8 warps (256 worlds) per SM if
    // the longest world seen so far then fits on chip, else with 4 warps (32 instead of 16 TMEM rows per world).
    const int seen = a->h_hw_rows && *(volatile int *)a->h_hw_rows > 0 ? *(volatile int *)a->h_hw_rows : maxrows;
    const int kr_t = a->rows_registers < 0 ? 8 : std::min(8, a->rows_registers & ~3);
    // warps per CTA: 8 (two per scheduler) unless the per-world arrays of 256 worlds leave no room for a pool
    int nw_t = a->sor_warps == 4 ? 4 : 8, pool_t = -1, want_t = 0;
    for (;; nw_t = 4) {
        const size_t fixed_t = b2p_sor_tmem_smem_bytes(nb, maxrows, 0, nw_t);
        pool_t = fixed_t + 1024 + 2 * 2048 <= kSmemMax ? (int)((kSmemMax - 1024 - fixed_t) / 2048) : -1;
        want_t = std::max(std::min(seen, maxrows) - b2p_sor_tmem_rows_on_chip(kr_t, nw_t), 0);
        if (nw_t == 4 || a->sor_warps == 8 || (pool_t >= 0 && want_t <= pool_t)) break;
    }
    // (the Tensor-Memory kernel addresses a world's row units with 32-bit byte offsets: row buffer < 4 GB)
    const bool rows_fit32 = (double)(maxrows + 1) * 64.0 * (double)a->Wp < 4294967296.0;
    use_tmem = a->sor_kernel == B2P_SOR_TMEM || (a->sor_kernel == B2P_SOR_AUTO && pool_t >= 0 && rows_fit32);
    if (use_tmem && (pool_t < 0 || !rows_fit32))
        return fail(B2P_ECAPACITY, "scene too large for the Tensor-Memory SOR kernel");
    if (use_tmem) {
        P.rows_registers = kr_t;
        P.sor_warps = nw_t;
        // one CTA per SM: spread the worlds evenly over whole waves of CTAs instead of leaving the last wave
        // partly empty (65536 worlds on 148 SMs: 2 waves of 224-world CTAs rather than 1.73 waves of 256)
        {
            const int tw = nw_t * 32, sms = std::max(a->sm_count, 1);
            const int waves = (a->W + tw * sms - 1) / (tw * sms);
            const int per = (a->W + waves * sms - 1) / (waves * sms);
            P.worlds_per_cta = std::min(tw, std::max(32, (per + 31) & ~31));
        }
        // Pool size: all the shared memory there is.  The kernel's row gathers do not allocate in L1 (LDG.NA), so
        // nothing is gained by leaving shared memory to the L1 carve-out (65536 rover worlds: 59 units 0.266 ms,
        // 48 units 0.292 ms; with L1-allocating gathers it was the other way round, 0.307 against 0.279 ms).
        const int pool_auto = pool_t;
        P.rows_resident = a->rows_resident > 0 ? std::max(std::min(a->rows_resident, pool_t), 2) : pool_auto;
    } else {
        P.rows_registers = b2p_sor_register_rows(maxrows, a->rows_registers);
        P.rows_resident = pick_rows_resident(a, nb, maxrows, P.rows_registers);
    }
    P.stepper = a->stepper;
    if (a->stepper == B2P_STEPPER_DENSE) {
        if (maxrows > 254) return fail(B2P_ECAPACITY, "the dense stepper supports at most 254 rows per world");
        // the per-world global scratch (2 x rows^2 doubles per world) only where the thread-per-world kernel runs
        const bool need_scratch = (P.tune & 8) != 0 || !b2p_dense_fits_warp_kernel(nb, std::max(maxrows, 1));
        if (need_scratch && (a->dense_rows != maxrows || a->dense_nb != nb || !a->dense.A)) {
            if (a->dense.A) cudaFree(a->dense.A);
            if (a->dense.L) cudaFree(a->dense.L);
            if (a->dense.vec) cudaFree(a->dense.vec);
            if (a->dense.sets) cudaFree(a->dense.sets);
            a->dense.A = a->dense.L = a->dense.vec = nullptr;
            a->dense.sets = nullptr;
            size_t md = 0, vd = 0, sb = 0;
            b2p_dense_scratch_sizes(nb, std::max(maxrows, 1), &md, &vd, &sb);
            CUDA_TRY(cudaMalloc(&a->dense.A, sizeof(double) * md * (size_t)a->Wp));
            CUDA_TRY(cudaMalloc(&a->dense.L, sizeof(double) * md * (size_t)a->Wp));
            CUDA_TRY(cudaMalloc(&a->dense.vec, sizeof(double) * vd * (size_t)a->Wp));
            CUDA_TRY(cudaMalloc(&a->dense.sets, sb * (size_t)a->Wp));
            a->dense_rows = maxrows;
            a->dense_nb = nb;
        }
        if (!a->dense.failures) {
            CUDA_TRY(cudaMalloc(&a->dense.failures, sizeof(int)));
            CUDA_TRY(cudaMemsetAsync(a->dense.failures, 0, sizeof(int), a->stream));
        }
        if (!a->dense.residual) CUDA_TRY(cudaMalloc(&a->dense.residual, sizeof(float) * (size_t)a->Wp));
    }
    if (a->order_mode == B2P_ORDER_ODE_ISLANDS && a->maxc > 64)
        return fail(B2P_ECAPACITY, "the island row order supports at most 64 contacts per world");
    if (b2p_assemble_smem_bytes(nb, (int)a->joints.size(), a->maxc, a->order_mode) > kSmemMax ||
        (!use_tmem && b2p_sor_smem_bytes(nb, maxrows, P.rows_resident) > kSmemMax))
        return fail(B2P_ECAPACITY, "scene needs more than %zu bytes of shared memory per CTA", kSmemMax);
    return 0;
}

// The assemble kernel is the only reader of the joint commands: an event behind it lets the next command upload
// start while the step's SOR-LCP kernel is still running (b2p_joints_set_param_batch).
cudaError_t mark_assemble_done(b2p_arena *a, cudaStream_t st)
{
    if (!a->ev_asm_done) {
        cudaError_t e = cudaEventCreateWithFlags(&a->ev_asm_done, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    a->asm_recorded = true;
    return cudaEventRecord(a->ev_asm_done, st);
}

// the three launches of a step on stream `st` (a real stream, or the capture stream of a graph); `part`: 0 the
// assemble kernel, 1 the SOR-LCP and integrate kernels, -1 all three
int enqueue_step(b2p_arena *a, const StepParams &P, bool use_tmem, cudaStream_t st, bool capturing, int part)
{
    const int nb = (int)a->bodies.size();
    const int maxrows = (int)a->joints.size() * B2P_ROWS_PER_JOINT + a->maxc * rpc_of(a);
    const bool timing = a->timing && !capturing;
    if (part != 1) {
        if (timing) CUDA_TRY(cudaEventRecord(a->ev[0], st));
        CUDA_TRY(b2p_launch_assemble(&P, nb, (int)a->joints.size(), a->maxc, st));
        if (!capturing) CUDA_TRY(mark_assemble_done(a, st));
        if (timing) CUDA_TRY(cudaEventRecord(a->ev[1], st));
    }
    if (part != 0) {
        if (P.stepper == B2P_STEPPER_DENSE) CUDA_TRY(b2p_launch_dense(&P, &a->dense, nb, maxrows, a->sm_count, (P.tune & 8) != 0, st));
        else if (use_tmem) CUDA_TRY(b2p_launch_sor_tmem(&P, nb, maxrows, st));
        else CUDA_TRY(b2p_launch_sor(&P, nb, maxrows, st));
        if (timing) CUDA_TRY(cudaEventRecord(a->ev[2], st));
        CUDA_TRY(b2p_launch_integrate(&P, nb, (int)a->joints.size(), st));
        if (timing) CUDA_TRY(cudaEventRecord(a->ev[3], st));
    }
    return 0;
}

void after_step(b2p_arena *a, const StepParams &P, bool use_tmem)
{
    a->sor_kernel_used = use_tmem ? B2P_SOR_TMEM : B2P_SOR_PAIRED;
    a->lastP = P;
    a->steps_done++;
    a->launches += 3;
    a->stepped = true;
    a->state.dev_newer = true;
    a->jobs.dev_newer = a->jfb.dev_newer = a->cfb.dev_newer = true;
    a->bodyx.dev_newer = a->bodyx.dev != nullptr;
    a->jupd.dev_newer = a->jupd.dev != nullptr;
    a->nrows.dev_newer = a->seed.dev_newer = a->jrows.dev_newer = true;
    // The step leaves the force accumulators on the device: zero (ODE: "zero all force accumulators") plus what
    // Joint::applyJointFriction added for the next step.  The host mirror is NOT touched (no O(worlds) host work
    // per step): the device copy is the current one, one-world accessors fetch / update their column.
    a->force.host_newer = false;
    a->force.dev_newer = true;
}

// Collide (optional) + step.  Launch shape `which` (0 = step, 1 = collide + step) is replayed from a CUDA graph
// when its parameter blocks equal the previous call's; the call in which they change launches directly (that is
// also where a kernel's shared-memory attribute gets configured) and the next identical one re-captures.
int run_step(b2p_arena *a, double step_size, bool with_collide)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    if (!(step_size > 0)) return fail(B2P_EINVAL, "step size must be positive");
    CUDA_TRY(cudaSetDevice(a->device));
    CollideParams C;
    memset(&C, 0, sizeof(C));
    int sampled = 0, rc;
    if (with_collide && (rc = prepare_collide(a, C, sampled))) return rc;
    StepParams P;
    bool use_tmem = false;
    if ((rc = prepare_step(a, step_size, P, use_tmem))) return rc;
    if (with_collide) { // finalize() inside prepare_step may have (re)allocated nothing new, but keep both views equal
        C.state = a->state.dev;
        C.crec = a->crec.dev;
        C.ccount = a->ccount.dev;
    }
    if ((rc = join_pending_upload(a))) return rc;
    const int ng = (int)a->geoms.size();
    b2p_arena::StepGraph &g = a->graphs[with_collide ? 1 : 0];
    const bool same = memcmp(&g.P, &P, sizeof(P)) == 0 && g.use_tmem == (int)use_tmem &&
                      (!with_collide || (memcmp(&g.C, &C, sizeof(C)) == 0 && g.sampled == sampled));
    const bool graph_ok = a->use_graph && !a->timing;
    if (graph_ok && same) {
        const int first = a->upload_mode ? 0 : 2, last = a->upload_mode ? 1 : 2;
        if (!g.exec[first]) {
            if (!a->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&a->cap_stream, cudaStreamNonBlocking));
            for (int gi = first; gi <= last; gi++) {
                const int part = gi == 2 ? -1 : gi;
                cudaGraph_t graph = nullptr;
                CUDA_TRY(cudaStreamBeginCapture(a->cap_stream, cudaStreamCaptureModeThreadLocal));
                int erc = 0;
                if (part != 1 && with_collide) {
                    cudaError_t e = b2p_launch_collide(&C, ng, sampled, a->cap_stream);
                    if (e != cudaSuccess) erc = fail(B2P_ECUDA, "collide launch (capture): %s", cudaGetErrorString(e));
                }
                if (!erc) erc = enqueue_step(a, P, use_tmem, a->cap_stream, true, part);
                cudaError_t ee = cudaStreamEndCapture(a->cap_stream, &graph);
                if (!erc && ee != cudaSuccess) erc = fail(B2P_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ee));
                if (!erc) {
                    ee = cudaGraphInstantiate(&g.exec[gi], graph, 0);
                    if (ee != cudaSuccess) {
                        g.exec[gi] = nullptr;
                        erc = fail(B2P_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ee));
                    }
                }
                if (graph) cudaGraphDestroy(graph);
                if (erc) {
                    for (int i = first; i <= last; i++) {
                        if (g.exec[i]) cudaGraphExecDestroy(g.exec[i]);
                        g.exec[i] = nullptr;
                    }
                    return erc;
                }
            }
        }
        if (a->upload_mode) {
            CUDA_TRY(cudaGraphLaunch(g.exec[0], a->stream));
            CUDA_TRY(mark_assemble_done(a, a->stream));
            CUDA_TRY(cudaGraphLaunch(g.exec[1], a->stream));
        } else {
            CUDA_TRY(cudaGraphLaunch(g.exec[2], a->stream));
        }
    } else {
        for (int i = 0; i < 3; i++) {
            if (g.exec[i]) cudaGraphExecDestroy(g.exec[i]);
            g.exec[i] = nullptr;
        }
        g.P = P;
        g.C = C;
        g.use_tmem = use_tmem;
        g.sampled = sampled;
        if (with_collide) CUDA_TRY(b2p_launch_collide(&C, ng, sampled, a->stream));
        if ((rc = enqueue_step(a, P, use_tmem, a->stream, false, -1))) return rc;
    }
    if (with_collide) after_collide(a);
    // refresh the host copy of the row high-water mark (every step at first, then occasionally)
    if (a->steps_done < 8 || (a->steps_done & 31) == 0)
        CUDA_TRY(cudaMemcpyAsync(a->h_hw_rows, a->d_hw_rows, sizeof(int), cudaMemcpyDeviceToHost, a->stream));
    if (a->timing) {
        CUDA_TRY(cudaEventSynchronize(a->ev[3]));
        for (int i = 0; i < 3; i++) {
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, a->ev[i], a->ev[i + 1]));
            a->kernel_ms[i] += ms;
        }
        a->timed_steps++;
    }
    after_step(a, P, use_tmem);
    return 0;
}
} // namespace

int b2p_collide(b2p_arena *a)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    CUDA_TRY(cudaSetDevice(a->device));
    CollideParams P;
    int sampled = 0;
    int rc = prepare_collide(a, P, sampled);
    if (rc) return rc;
    CUDA_TRY(b2p_launch_collide(&P, (int)a->geoms.size(), sampled, a->stream));
    after_collide(a);
    return 0;
}

int b2p_collide_record_pairs(b2p_arena *a, int on)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->want_pairs = on != 0;
    return 0;
}

int b2p_contacts_dropped(b2p_arena *a, long long *count)
{
    if (!a || !count) return fail(B2P_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(a->device));
    *count = 0;
    if (!a->d_dropped) return 0;
    int n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, a->d_dropped, sizeof(int), cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    *count = n;
    return 0;
}

int b2p_collide_get_pairs(b2p_arena *a, int world, int *pairs, int max_pairs)
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    if (!a->want_pairs) return fail(B2P_EINVAL, "call b2p_collide_record_pairs(a, 1) before b2p_collide");
    if (!a->d_pairs) return 0;
    std::vector<int> col((size_t)a->maxpairs);
    int n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, a->d_npairs + world, sizeof(int), cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaMemcpy2DAsync(col.data(), sizeof(int), a->d_pairs + world, sizeof(int) * a->Wp, sizeof(int),
                               (size_t)a->maxpairs, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    if (n > a->maxpairs) n = a->maxpairs;
    for (int i = 0; i < n && i < max_pairs; i++) pairs[i] = col[(size_t)i];
    return n;
}

// ---------------- the step ----------------
int b2p_step(b2p_arena *a, double step_size) { return run_step(a, step_size, false); }

// b2p_collide + b2p_step as one call (and, from the second identical call on, one CUDA graph launch of the four
// kernels): the per-tick sequence of a device-collision scene
int b2p_collide_step(b2p_arena *a, double step_size)
{
    if (a && a->geoms.empty()) return run_step(a, step_size, false);
    return run_step(a, step_size, true);
}

int b2p_arena_set_tune(b2p_arena *a, int bits)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->tune = bits;
    return 0;
}

int b2p_arena_set_graph_launch(b2p_arena *a, int on)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->use_graph = on != 0;
    return 0;
}

int b2p_sync(b2p_arena *a)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

int b2p_world_last_num_rows(b2p_arena *a, int world, int *rows)
{
    int rc = check_world(a, world, false);
    if (rc) return rc;
    *rows = 0;
    if (a->nrows.fields == 0) return 0;
    if ((rc = mirror_fetch_column(a, a->nrows, 0, 1, world))) return rc;
    *rows = a->nrows.host[(size_t)world] & 0xffff; // high half: independent rows (device use)
    return 0;
}

int b2p_kernel_launch_count(const b2p_arena *a) { return a ? a->launches : 0; }

int b2p_arena_set_resident_rows(b2p_arena *a, int n)
{
    if (!a || n < 0) return fail(B2P_EINVAL, "bad resident row count");
    a->rows_resident = n;
    return 0;
}

int b2p_arena_set_sor_kernel(b2p_arena *a, int which)
{
    if (!a || which < B2P_SOR_AUTO || which > B2P_SOR_TMEM) return fail(B2P_EINVAL, "bad SOR kernel selector");
    a->sor_kernel = which;
    return 0;
}

int b2p_arena_sor_kernel_used(const b2p_arena *a) { return a ? a->sor_kernel_used : 0; }

int b2p_arena_set_sor_warps(b2p_arena *a, int warps)
{
    if (!a || (warps != 0 && warps != 4 && warps != 8)) return fail(B2P_EINVAL, "warps per CTA must be 0, 4 or 8");
    a->sor_warps = warps;
    return 0;
}

int b2p_arena_set_stepper(b2p_arena *a, int stepper)
{
    if (!a || (stepper != B2P_STEPPER_QUICK && stepper != B2P_STEPPER_DENSE)) return fail(B2P_EINVAL, "bad stepper selector");
    a->stepper = stepper;
    return 0;
}

int b2p_arena_stepper(const b2p_arena *a) { return a ? a->stepper : 0; }

int b2p_dense_failures(b2p_arena *a, long long *count)
{
    if (!a || !count) return fail(B2P_EINVAL, "null argument");
    *count = 0;
    if (!a->dense.failures) return 0;
    CUDA_TRY(cudaSetDevice(a->device));
    int v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, a->dense.failures, sizeof(int), cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    *count = v;
    return 0;
}

int b2p_dense_residual(b2p_arena *a, float *host_out)
{
    if (!a || !host_out) return fail(B2P_EINVAL, "null argument");
    if (!a->stepped || !a->dense.residual || a->stepper != B2P_STEPPER_DENSE) return fail(B2P_EINVAL, "no dense step has run yet");
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaMemcpyAsync(host_out, a->dense.residual, sizeof(float) * (size_t)a->W, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

int b2p_arena_set_register_rows(b2p_arena *a, int n)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->rows_registers = n;
    return 0;
}

// diagnostics: enable (and reset) per-kernel timing of b2p_step; while enabled every step synchronises.
// out[0..2] = mean milliseconds of the assemble / SOR / integrate kernels, out[3] = steps accumulated.
int b2p_debug_kernel_times(b2p_arena *a, int enable, double out[4])
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    CUDA_TRY(cudaSetDevice(a->device));
    if (out) {
        for (int i = 0; i < 3; i++) out[i] = a->timed_steps ? a->kernel_ms[i] / (double)a->timed_steps : 0.0;
        out[3] = (double)a->timed_steps;
    }
    a->timing = enable != 0;
    a->timed_steps = 0;
    a->kernel_ms[0] = a->kernel_ms[1] = a->kernel_ms[2] = 0;
    if (enable)
        for (cudaEvent_t &e : a->ev)
            if (!e) CUDA_TRY(cudaEventCreate(&e));
    return 0;
}

// page-locked host memory for the batched transfers (b2p_joints_set_param_batch, b2p_observe_begin ...): a host
// written in C++ / C has no CUDA runtime of its own to get it from
void *b2p_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        fail(B2P_ECUDA, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void b2p_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---------------- post-step bookkeeping of WorldPhysics::stepTheWorld, for every world ----------------
int b2p_arena_set_post_step_outputs(b2p_arena *a, int flags)
{
    if (!a) return fail(B2P_EINVAL, "null arena");
    a->post_flags = flags & (B2P_POST_JOINT_UPDATE | B2P_POST_FRAME_UPDATE);
    return 0;
}

int b2p_body_set_damping(b2p_arena *a, int body, double linear, double angular)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    a->bodies[body].lin_damp = linear;
    a->bodies[body].ang_damp = angular;
    a->tmpl_dirty = true;
    return 0;
}

int b2p_body_get_damping(b2p_arena *a, int body, double *linear, double *angular)
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if (linear) *linear = a->bodies[body].lin_damp;
    if (angular) *angular = a->bodies[body].ang_damp;
    return 0;
}

int b2p_joint_set_friction_coefficient(b2p_arena *a, int ji, double k)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    a->joints[ji].friction = k;
    a->tmpl_dirty = true;
    return 0;
}

int b2p_joint_get_friction_coefficient(b2p_arena *a, int ji, double *k)
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if (k) *k = a->joints[ji].friction;
    return 0;
}

int b2p_joint_get_update(b2p_arena *a, int ji, int world, double out[7])
{
    int rc = check_joint(a, ji);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    for (int i = 0; i < 7; i++) out[i] = 0;
    if (a->jupd.fields < (size_t)(ji + 1) * 7) return 0; // never stepped, or the output is switched off
    return body_vec(a, a->jupd, (size_t)ji * 7, world, 7, out);
}

int b2p_body_get_acceleration(b2p_arena *a, int body, int world, double lin[3], double ang[3])
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (a->bodyx.fields >= (size_t)(body + 1) * 15 && (rc = body_vec(a, a->bodyx, (size_t)body * 15 + 6, world, 6, v)))
        return rc;
    for (int i = 0; i < 3; i++) {
        if (lin) lin[i] = v[i];
        if (ang) ang[i] = v[3 + i];
    }
    return 0;
}

int b2p_body_get_contact_force(b2p_arena *a, int body, int world, double out[3])
{
    int rc = check_body(a, body);
    if (rc) return rc;
    if ((rc = check_world(a, world, false))) return rc;
    out[0] = out[1] = out[2] = 0;
    if (a->bodyx.fields < (size_t)(body + 1) * 15) return 0;
    return body_vec(a, a->bodyx, (size_t)body * 15 + 12, world, 3, out);
}

static int download_fields(b2p_arena *a, const float *dev, size_t fields, float *host)
{
    if (!a || !host) return fail(B2P_EINVAL, "null argument");
    if (!a->stepped || !dev) return fail(B2P_EINVAL, "no step has run yet (or the output is switched off)");
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaMemcpy2DAsync(host, sizeof(float) * a->W, dev, sizeof(float) * a->Wp, sizeof(float) * a->W, fields,
                               cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}
int b2p_joint_update_download(b2p_arena *a, float *host)
{
    return download_fields(a, a ? a->jupd.dev : nullptr, a ? a->joints.size() * 7 : 0, host);
}
int b2p_body_post_download(b2p_arena *a, float *host)
{
    return download_fields(a, a ? a->bodyx.dev : nullptr, a ? a->bodies.size() * 15 : 0, host);
}
void *b2p_device_joint_update(b2p_arena *a) { return a ? a->jupd.dev : nullptr; }
void *b2p_device_body_post(b2p_arena *a) { return a ? a->bodyx.dev : nullptr; }

// LCP residual of the last step's solve, one value per world (see b2p_residual_kernel)
int b2p_debug_lcp_residual(b2p_arena *a, float *host_out)
{
    if (!a || !host_out) return fail(B2P_EINVAL, "null argument");
    if (!a->stepped) return fail(B2P_EINVAL, "no step has run yet");
    CUDA_TRY(cudaSetDevice(a->device));
    if (!a->d_residual) CUDA_TRY(cudaMalloc(&a->d_residual, sizeof(float) * (size_t)a->Wp));
    CUDA_TRY(b2p_launch_residual(&a->lastP, a->d_residual, a->stream));
    a->launches++;
    CUDA_TRY(cudaMemcpyAsync(host_out, a->d_residual, sizeof(float) * (size_t)a->W, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return 0;
}

// body 212 B, hinge 160 B, fixed 88 B, slider 132 B, hinge2 212 B,
// contact 76 B (+128 B when the record was produced on the device), geom 48 B.
double b2p_algorithmic_bytes_per_world_step(const b2p_arena *ca)
{
    b2p_arena *a = const_cast<b2p_arena *>(ca);
    if (!a) return 0;
    double bytes = 212.0 * a->bodies.size();
    for (const HJoint &j : a->joints) {
        switch (j.type) {
        case B2P_JOINT_HINGE: bytes += 160; break;
        case B2P_JOINT_FIXED: bytes += 88; break;
        case B2P_JOINT_SLIDER: bytes += 132; break;
        case B2P_JOINT_HINGE2: bytes += 212; break;
        case B2P_JOINT_BALL: bytes += 88; break;       // ids, anchors 6, erp, cfm in; feedback 12 out (as fixed)
        case B2P_JOINT_UNIVERSAL: bytes += 228; break; // hinge2's 212 + qrel2
        default: break;
        }
    }
    double mean_contacts = 0;
    if (a->ccount.fields && mirror_to_host(a, a->ccount) == 0) {
        for (int w = 0; w < a->W; w++) mean_contacts += a->ccount.host[(size_t)w];
        mean_contacts /= a->W;
    }
    const bool device_contacts = !a->geoms.empty();
    bytes += mean_contacts * (76.0 + (device_contacts ? 128.0 : 0.0));
    bytes += 48.0 * a->geoms.size();
    return bytes;
}

} // extern "C"
