// GenSimDefs.h -- containers of the host layer.  The reference typedefs blitz::Array<double,2/3> here
// (src/GenSimDefs.h:25-31); blitz++ is not a dependency of this build, so Array2d / Array3d are small own classes
// with the blitz surface the solver code uses (resize, reindexSelf, operator(), = scalar, copy, base, extent).
//
// B200 twist: an array can be BOUND to a field of a device context (impec_ctx).  The device copy is then the
// working copy; host element access triggers a lazy download, host writes mark the host copy dirty and the next
// device operation uploads it.  A time loop that never touches elements on the host never moves a byte over PCIe.
#ifndef ARRAYTYPEDEFS_H_
#define ARRAYTYPEDEFS_H_
#include <vector>
#include <cstddef>
#define SPLIT2ND                       // second-order operator splitting, as upstream (src/GenSimDefs.h:18)

struct impec_ctx;
enum DeviceField { FIELD_NONE = -1, FIELD_C = 0, FIELD_P = 1, FIELD_V = 2 };

template <int R> struct IdxN {
    int v[R];
    int &operator[](int k) { return v[k]; }
    int operator[](int k) const { return v[k]; }
};
typedef IdxN<2> Idx2d;
typedef IdxN<3> Idx3d;

class DeviceBound {
protected:
    mutable impec_ctx *ctx_ = nullptr;
    mutable int field_ = FIELD_NONE;
    mutable bool host_dirty_ = true;   // host copy modified since the last upload
    mutable bool dev_newer_ = false;   // device copy modified since the last download
    virtual double *raw() const = 0;
    void pull() const;                 // download if the device copy is newer
public:
    DeviceBound() {}
    // a copy is a plain host array: the source is brought up to date first, the binding is NOT inherited
    DeviceBound(const DeviceBound &o) { o.pull(); }
    DeviceBound &operator=(const DeviceBound &o) { if (this != &o) { o.pull(); host_dirty_ = true; dev_newer_ = false; } return *this; }
    virtual ~DeviceBound();                       // leaves the registry of bound arrays (DeviceContext.cpp)
    void Bind(impec_ctx *ctx, int field) const;   // (re)bind; a new binding starts with the host copy as truth
    void Unbind() const;                          // download if the device copy is newer, then detach
    static void UnbindAll(const impec_ctx *ctx);  // detach every array bound to ctx (called before ctx is destroyed)
    void SyncToDevice() const;         // upload if the host copy is dirty
    void MarkDeviceNewer() const { dev_newer_ = true; host_dirty_ = false; }
    void GatherGlobal(const char *field) const;   // multi-GPU: fill the host copy with the rows of ALL ranks (collective; impec_gather_rows)
    bool BoundTo(const impec_ctx *c) const { return ctx_ == c; }
};

class Array2d : public DeviceBound {
    std::vector<double> d_;
    int b_[2] = {0, 0}, e_[2] = {0, 0};
    double *raw() const override { return const_cast<double *>(d_.data()); }
public:
    Array2d() {}
    Array2d(int n0, int n1) { resize(n0, n1); }
    void resize(int n0, int n1) { e_[0] = n0; e_[1] = n1; d_.assign((size_t)n0 * n1, 0.0); host_dirty_ = true; dev_newer_ = false; }
    void reindexSelf(const Idx2d &b) { b_[0] = b[0]; b_[1] = b[1]; }
    double &operator()(int i, int j) { pull(); host_dirty_ = true; return d_[(size_t)(i - b_[0]) * e_[1] + (j - b_[1])]; }
    double operator()(int i, int j) const { pull(); return d_[(size_t)(i - b_[0]) * e_[1] + (j - b_[1])]; }
    Array2d &operator=(double v) { d_.assign(d_.size(), v); host_dirty_ = true; dev_newer_ = false; return *this; }
    const int *base() const { return b_; }
    const int *extent() const { return e_; }
    int extent(int k) const { return e_[k]; }
    double *data() { pull(); host_dirty_ = true; return d_.data(); }
    const double *data() const { pull(); return d_.data(); }
    size_t size() const { return d_.size(); }
    Array2d copy() const { pull(); Array2d r; r.d_ = d_; r.b_[0] = b_[0]; r.b_[1] = b_[1]; r.e_[0] = e_[0]; r.e_[1] = e_[1]; return r; }
};

class Array3d : public DeviceBound {
    std::vector<double> d_;
    int b_[3] = {0, 0, 0}, e_[3] = {0, 0, 0};
    double *raw() const override { return const_cast<double *>(d_.data()); }
public:
    void resize(int n0, int n1, int n2) { e_[0] = n0; e_[1] = n1; e_[2] = n2; d_.assign((size_t)n0 * n1 * n2, 0.0); host_dirty_ = true; dev_newer_ = false; }
    void reindexSelf(const Idx3d &b) { b_[0] = b[0]; b_[1] = b[1]; b_[2] = b[2]; }
    double &operator()(int i, int j, int k) { pull(); host_dirty_ = true; return d_[((size_t)(i - b_[0]) * e_[1] + (j - b_[1])) * e_[2] + (k - b_[2])]; }
    double operator()(int i, int j, int k) const { pull(); return d_[((size_t)(i - b_[0]) * e_[1] + (j - b_[1])) * e_[2] + (k - b_[2])]; }
    Array3d &operator=(double v) { d_.assign(d_.size(), v); host_dirty_ = true; dev_newer_ = false; return *this; }
    const int *base() const { return b_; }
    const int *extent() const { return e_; }
    double *data() { pull(); host_dirty_ = true; return d_.data(); }
    const double *data() const { pull(); return d_.data(); }
};
#endif
