#include "DeviceContext.h"
#include <map>
#include <set>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <fstream>
#include <string>
#include <chrono>
#include <thread>
#include <cstdio>
#include <unistd.h>

namespace {
int g_rank = 0, g_nranks = 1;
// NCCL unique id through the file system: rank 0 writes <file>.tmp and renames it into place, the other ranks poll for it
std::string id_file() {
    if (const char *e = getenv("IMPEC_NCCL_ID_FILE")) return e;
    const char *port = getenv("MASTER_PORT");
    return std::string("/tmp/impec2d_nccl_id.") + (port ? std::string(port) : std::to_string((long)getppid()));
}
bool exchange_nccl_id(char id[128]) {
    const std::string path = id_file();
    if (g_rank == 0) {
        if (impec_nccl_unique_id(id) != IMPEC_OK) return false;
        const std::string tmp = path + ".tmp";
        { std::ofstream f(tmp.c_str(), std::ios::binary); if (!f.write(id, 128)) return false; }
        return std::rename(tmp.c_str(), path.c_str()) == 0;
    }
    for (int tries = 0; tries < 1200; tries++) {                 // up to 60 s
        std::ifstream f(path.c_str(), std::ios::binary);
        if (f && f.read(id, 128)) return true;
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    return false;
}
struct Entry { impec_ctx *ctx = nullptr; impec_params P; const double *q_uploaded = nullptr; };
std::map<SimParams *, Entry> g_ctx;
// arrays currently bound to a context: a context is never destroyed under them (DeviceBound::UnbindAll)
std::set<const DeviceBound *> &bound_set() { static std::set<const DeviceBound *> s; return s; }
}

void DeviceContext::Check(impec_ctx *ctx, int rc, const char *what) {
    if (rc == IMPEC_OK) return;
    if (rc > 0) {                                        // warning (a solver stopped at max_iter): upstream never checks, we say so and go on
        std::cerr << "impec2d_b200: warning from " << what << ": " << impec_last_error(ctx) << std::endl;
        return;
    }
    std::cerr << "impec2d_b200: " << what << " failed (" << rc << "): " << impec_last_error(ctx) << std::endl;
    exit(1);                                             // the reference's error convention for fatal errors
}

impec_params DeviceContext::ToParams(const SimParams *sp) {
    impec_params P; impec_default_params(&P);
    P.N1 = sp->N1; P.N2 = sp->N2; P.g = sp->c_bounds;
    P.x_lb = sp->x_lb; P.x_ub = sp->x_ub; P.y_lb = sp->y_lb; P.y_ub = sp->y_ub;
    P.mu_res = sp->mu_res; P.mu_inj = sp->mu_inj;
    P.avgproc = sp->avgproc == HARMONIC ? IMPEC_HARMONIC : IMPEC_MEANCONC;
    P.nu = sp->nu; P.is_test = sp->is_test ? 1 : 0; P.testno = sp->is_test ? sp->testno : 0;
    P.stype = sp->stype == PARALLEL ? IMPEC_PARALLEL : (sp->stype == DIAGONAL ? IMPEC_DIAGONAL : IMPEC_HORIZONTAL);
    P.porosity = sp->porosity; P.cfl = sp->cfl; P.inj_fluid_conc = sp->inj_fluid_conc;
    P.diffusion = sp->diffusion ? 1 : 0; P.alpha_l = sp->alpha_l; P.alpha_t = sp->alpha_t; P.dt_diff_limit = sp->dt_diff_limit;
    P.ptype = sp->ptype == CRANKNICOLSON ? IMPEC_CRANKNICOLSON : IMPEC_EXPLICIT;
    P.eps = sp->eps; P.para_eps = sp->para_eps;
    // hyper_method: the reference reads the key but hard-wires MUSCLSplitSolver (src/Impec2DSolver.cpp:132); here the method
    // belongs to the solver OBJECT and is passed per call (DeviceContext::SetMethod), so it is not part of the context identity
    P.hyper_method = IMPEC_MUSCL;
    switch (sp->dim_splitting) {
    case ALTERNATING: P.dim_splitting = IMPEC_ALTERNATING; break;
    case X_FIRST: P.dim_splitting = IMPEC_X_FIRST; break;
    case Y_FIRST: P.dim_splitting = IMPEC_Y_FIRST; break;
    case STRANG: P.dim_splitting = IMPEC_STRANG; break;
    default: P.dim_splitting = IMPEC_SYMMETRIC_PARALLEL; break;
    }
    return P;
}

impec_ctx *DeviceContext::Get(SimParams *sp) {
    Entry &e = g_ctx[sp];
    const impec_params P = ToParams(sp);
    if (e.ctx && std::memcmp(&P, &e.P, sizeof(P)) == 0) return e.ctx;
    if (e.ctx) { DeviceBound::UnbindAll(e.ctx); impec_ctx_destroy(e.ctx); e.ctx = nullptr; }   // grid/physics changed: arrays keep their data on the host
    const char *dev = getenv("IMPEC_DEVICE"), *lr = getenv("LOCAL_RANK");
    const int device = dev ? atoi(dev) : (g_nranks > 1 ? (lr ? atoi(lr) : g_rank) : 0);
    char id[128]; const char *idp = nullptr;
    if (g_nranks > 1 && !impec_is_emulated()) {              // (the emulated test build connects its processes through IMPEC_EMU_COMM_DIR instead)
        if (!exchange_nccl_id(id)) { std::cerr << "impec2d_b200: rank " << g_rank << " could not exchange the NCCL id through " << id_file() << std::endl; exit(1); }
        idp = id;
    }
    const int rc = impec_ctx_create(&P, device, g_rank, g_nranks, idp, nullptr, &e.ctx);
    Check(e.ctx, rc, "impec_ctx_create");
    if (g_nranks > 1 && g_rank == 0 && !impec_is_emulated()) std::remove(id_file().c_str());     // every rank has joined the communicator
    e.P = P; e.q_uploaded = nullptr;
    return e.ctx;
}

void DeviceContext::Release(SimParams *sp) {
    auto it = g_ctx.find(sp);
    if (it == g_ctx.end()) return;
    if (it->second.ctx) { DeviceBound::UnbindAll(it->second.ctx); impec_ctx_destroy(it->second.ctx); }
    g_ctx.erase(it);
}

void DeviceContext::SetRanks(int rank, int nranks) { g_rank = rank; g_nranks = nranks < 1 ? 1 : nranks; }
int DeviceContext::Rank() { return g_rank; }
int DeviceContext::NumRanks() { return g_nranks; }
void DeviceContext::GatherGlobal(impec_ctx *ctx, const Array2d *c, const Array2d *p, const Array3d *v) {
    if (c) { if (!c->BoundTo(ctx)) c->Bind(ctx, FIELD_C); c->GatherGlobal("c"); }
    if (p) { if (!p->BoundTo(ctx)) p->Bind(ctx, FIELD_P); p->GatherGlobal("p"); }
    if (v) { if (!v->BoundTo(ctx)) v->Bind(ctx, FIELD_V); v->GatherGlobal("v"); }
}

void DeviceContext::SetMethod(impec_ctx *ctx, int method) { Check(ctx, impec_set_option(ctx, "hyper_method", (double)method), "impec_set_option(hyper_method)"); }

void DeviceContext::InvalidateSource(SimParams *sp) { g_ctx[sp].q_uploaded = nullptr; }

void DeviceContext::Use(impec_ctx *ctx, const Array2d *c, const Array2d *p, const Array3d *v, const double *q, SimParams *sp) {
    if (c) { if (!c->BoundTo(ctx)) c->Bind(ctx, FIELD_C); c->SyncToDevice(); }
    if (p) { if (!p->BoundTo(ctx)) p->Bind(ctx, FIELD_P); p->SyncToDevice(); }
    if (v) { if (!v->BoundTo(ctx)) v->Bind(ctx, FIELD_V); v->SyncToDevice(); }
    if (q) {
        Entry &e = g_ctx[sp];
        if (e.q_uploaded != q) { Check(ctx, impec_upload_q(ctx, q), "impec_upload_q"); e.q_uploaded = q; }
    }
}

// ---- DeviceBound (declared in GenSimDefs.h)
void DeviceBound::Bind(impec_ctx *ctx, int field) const {
    if (ctx_ && ctx_ != ctx) pull();
    ctx_ = ctx; field_ = field; host_dirty_ = true; dev_newer_ = false;
    bound_set().insert(this);
}
void DeviceBound::Unbind() const { pull(); ctx_ = nullptr; field_ = FIELD_NONE; bound_set().erase(this); }
DeviceBound::~DeviceBound() { bound_set().erase(this); }
void DeviceBound::UnbindAll(const impec_ctx *ctx) {
    std::vector<const DeviceBound *> mine;
    for (const DeviceBound *b : bound_set()) if (b->ctx_ == ctx) mine.push_back(b);
    for (const DeviceBound *b : mine) b->Unbind();
}
void DeviceBound::pull() const {
    if (!dev_newer_ || !ctx_) return;
    int rc = IMPEC_OK;
    if (field_ == FIELD_C) rc = impec_download_c(ctx_, raw());
    else if (field_ == FIELD_P) rc = impec_download_p(ctx_, raw());
    else if (field_ == FIELD_V) rc = impec_download_velocity(ctx_, raw());
    DeviceContext::Check(ctx_, rc, "download");
    dev_newer_ = false;
}
void DeviceBound::GatherGlobal(const char *field) const {
    if (!ctx_) return;
    SyncToDevice();                                  // host edits (if any) go up first: the gather reads the device copy
    DeviceContext::Check(ctx_, impec_gather_rows(ctx_, field, raw()), "impec_gather_rows");
    dev_newer_ = false;                              // the host copy now holds every rank's rows (own rows included)
}
void DeviceBound::SyncToDevice() const {
    if (!ctx_ || !host_dirty_) return;
    int rc = IMPEC_OK;
    if (field_ == FIELD_C) rc = impec_upload_c(ctx_, raw());
    else if (field_ == FIELD_P) rc = impec_upload_p(ctx_, raw());
    else if (field_ == FIELD_V) rc = impec_upload_velocity(ctx_, raw());
    DeviceContext::Check(ctx_, rc, "upload");
    host_dirty_ = false;
}
