// comm_p2p.inl -- halo rows and small reductions over NVLink peer memory (one process per GPU, one node).
//
// NCCL send/recv costs ~13-20 us per halo exchange on the solver's stream, and a multigrid cycle on slabs issues
// about 25 of them.  Here every rank owns a MAILBOX (one cudaMalloc, exported with cudaIpcGetMemHandle and mapped
// by the other ranks); an exchange is ONE small kernel that
//   1. stores this rank's boundary rows straight into the neighbours' mailboxes (128-bit stores over NVLink),
//   2. publishes a per-CTA flag (system-scope release) carrying the exchange's epoch,
//   3. waits for the matching flags of the neighbours (system-scope acquire), and
//   4. copies the received rows from its own mailbox into its halo rows.
// CTA b of the producer fills chunk b and raises flag b; CTA b of the consumer waits for flag b only, so no
// grid-wide barrier is needed.  Mailbox slots alternate with the parity of the epoch: a slot written in exchange
// e is unpacked by the consumer's kernel e, which has completed before the consumer's kernel e+1 raises the flag
// this rank waits for in ITS kernel e+1 -- so the slot is free again when kernel e+2 overwrites it.
// The all-reduce of a few scalars works the same way (every rank stores its contribution into every mailbox and
// combines the contributions in rank order, which makes the result bitwise identical on all ranks).
// All ranks must issue the same sequence of exchanges (they do: the solver is SPMD).  A wait that exceeds
// P2P_TIMEOUT_CYCLES raises a device-side error flag instead of hanging the GPU.
//
// The all-gather of the restricted right-hand side of the first replicated multigrid level (4 per cycle) is a kernel of the same kind
// (k_allgather_p2p): every rank stores its slab into a window of EVERY rank's mailbox, raises one flag per chunk and destination,
// waits for the matching chunks of all sources and copies the window into the level's array.  Two windows alternate with the epoch
// parity (a rank may be a whole tail visit ahead of a far-away rank; it can never be two gathers ahead, because completing gather e
// requires every rank to have posted gather e).
//
// NCCL remains the bootstrap (handle all-gather) and the transport for the once-per-step coefficient all-gathers.
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
#define P2P_MAXR 16
#define P2P_NCTA 8
#define P2P_THREADS 512
#define P2P_FLAG_BYTES 8192
#define P2P_AG_NCTA 32                       /* CTAs (= chunks per source slab) of the all-gather kernel */
#define P2P_AG_FLAG0 128                      /* first all-gather flag: flags [128 + src * P2P_AG_NCTA + cta] */
#define P2P_AR_BYTES 8192
#define P2P_AR_MAXN 8
#define P2P_TIMEOUT_CYCLES 120000000000LL   /* ~60 s: ranks may reach their first exchange seconds apart */

__device__ __forceinline__ void p2p_st_flag(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long p2p_ld_flag(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool p2p_wait(const unsigned long long *flag, unsigned long long epoch, double *err) {
    const long long t0 = clock64();
    while (p2p_ld_flag(flag) < epoch)
        if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { err[0] = 1.0; return false; }
    return true;
}

struct HaloArgs {
    double *fld[2]; int nfld;
    long long cnt;                                   // doubles per field and side (rows * pitch; a multiple of 2)
    long long send_lo, recv_lo, send_hi, recv_hi;    // element offsets inside a field
    double *rem_lo, *rem_hi;                         // slots in the previous / next rank's mailbox (nullptr at a wall)
    unsigned long long *rem_lo_flag, *rem_hi_flag;   // [P2P_NCTA] flags next to those slots
    const double *my_lo, *my_hi;                     // my slots: rows arriving from the previous / next rank
    const unsigned long long *my_lo_flag, *my_hi_flag;
    unsigned long long epoch;
    double *err;
};

__global__ void __launch_bounds__(P2P_THREADS) k_halo_p2p(HaloArgs a) {
    if (a.err[0] != 0.) return;                                            // an earlier exchange timed out: do not wait again
    const long long n2 = a.cnt / 2;                                        // double2 elements per field and side
    const long long chunk = (n2 + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * chunk, c1 = (c0 + chunk < n2) ? c0 + chunk : n2;
    // 1. put
    for (int f = 0; f < a.nfld; f++) {
        if (a.rem_lo) {
            const double2 *src = reinterpret_cast<const double2 *>(a.fld[f] + a.send_lo);
            double2 *dst = reinterpret_cast<double2 *>(a.rem_lo + (long long)f * a.cnt);
            for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) dst[k] = src[k];
        }
        if (a.rem_hi) {
            const double2 *src = reinterpret_cast<const double2 *>(a.fld[f] + a.send_hi);
            double2 *dst = reinterpret_cast<double2 *>(a.rem_hi + (long long)f * a.cnt);
            for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) dst[k] = src[k];
        }
    }
    __syncthreads();                       // the CTA's stores happen-before thread 0's release below (cumulativity)
    // 2. publish, 3. wait
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    // warp 0 talks to the previous rank, warp 1 to the next one (the two waits overlap)
    if (threadIdx.x == 0 && a.rem_lo) {
        p2p_st_flag(a.rem_lo_flag + blockIdx.x, a.epoch);
        if (!p2p_wait(a.my_lo_flag + blockIdx.x, a.epoch, a.err)) bad = 1;
    }
    if (threadIdx.x == 32 && a.rem_hi) {
        p2p_st_flag(a.rem_hi_flag + blockIdx.x, a.epoch);
        if (!p2p_wait(a.my_hi_flag + blockIdx.x, a.epoch, a.err)) bad = 1;
    }
    __syncthreads();                       // the acquire loads above order the unpack below after the neighbours' stores
    const int ok = !bad;
    if (!ok) return;
    // 4. unpack (L2 loads: the rows were written by another GPU)
    for (int f = 0; f < a.nfld; f++) {
        if (a.rem_lo) {
            const double2 *src = reinterpret_cast<const double2 *>(a.my_lo + (long long)f * a.cnt);
            double2 *dst = reinterpret_cast<double2 *>(a.fld[f] + a.recv_lo);
            for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) dst[k] = __ldcg(src + k);
        }
        if (a.rem_hi) {
            const double2 *src = reinterpret_cast<const double2 *>(a.my_hi + (long long)f * a.cnt);
            double2 *dst = reinterpret_cast<double2 *>(a.fld[f] + a.recv_hi);
            for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) dst[k] = __ldcg(src + k);
        }
    }
}

struct GatherArgs {
    const double *src;                               // my slab (cnt doubles, contiguous rows)
    long long cnt;                                   // doubles per rank (a multiple of 2)
    double *rem_win[P2P_MAXR];                       // this epoch's gather window in every rank's mailbox (rem_win[rank] = mine)
    unsigned long long *rem_flag[P2P_MAXR];          // flag area of every rank's mailbox
    double *dst;                                     // my replicated array at the offset of rank 0's slab
    int rank, nranks;
    unsigned long long epoch;
    double *err;
};
__global__ void __launch_bounds__(P2P_THREADS) k_allgather_p2p(GatherArgs a) {
    if (a.err[0] != 0.) return;
    const long long n2 = a.cnt / 2;
    const long long chunk = (n2 + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * chunk, c1 = (c0 + chunk < n2) ? c0 + chunk : n2;
    // 1. put my chunk into every rank's window (my own included: one code path, and the unpack below is uniform)
    const double2 *src = reinterpret_cast<const double2 *>(a.src);
    for (int r = 0; r < a.nranks; r++) {
        double2 *dst = reinterpret_cast<double2 *>(a.rem_win[(a.rank + r) % a.nranks] + (long long)a.rank * a.cnt);      // start with myself, then the ring: spreads the links
        for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    // 2. publish chunk blockIdx.x to every destination, 3. wait for chunk blockIdx.x of every source
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    if ((int)threadIdx.x < a.nranks) {
        const int t = threadIdx.x;
        p2p_st_flag(a.rem_flag[t] + P2P_AG_FLAG0 + a.rank * P2P_AG_NCTA + blockIdx.x, a.epoch);
        if (!p2p_wait(a.rem_flag[a.rank] + P2P_AG_FLAG0 + t * P2P_AG_NCTA + blockIdx.x, a.epoch, a.err)) bad = 1;
    }
    __syncthreads();
    if (bad) return;
    // 4. unpack: window -> the level's array (rows of all ranks are contiguous in both)
    const double *win = a.rem_win[a.rank];
    for (int s = 0; s < a.nranks; s++) {
        const double2 *w2 = reinterpret_cast<const double2 *>(win + (long long)s * a.cnt);
        double2 *d2p = reinterpret_cast<double2 *>(a.dst + (long long)s * a.cnt);
        for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) d2p[k] = __ldcg(w2 + k);
    }
}

// all-reduce of n <= P2P_AR_MAXN doubles, in place.  Mailbox record of source rank s, parity p:
//   doubles [ (p*P2P_MAXR + s) * (P2P_AR_MAXN+1) ... ) : n values, then the epoch flag in the last slot
struct ArArgs {
    double *buf; int n, op, rank, nranks;
    double *mail[P2P_MAXR];                           // all-reduce areas of every rank's mailbox (mail[rank] = mine)
    unsigned long long epoch;
    double *err;
};
__global__ void __launch_bounds__(32) k_allreduce_p2p(ArArgs a) {
    const int t = threadIdx.x;
    if (a.err[0] != 0.) return;
    const long long rec = (long long)((a.epoch & 1) * P2P_MAXR) * (P2P_AR_MAXN + 1);
    __shared__ int ok;
    if (t == 0) ok = 1;
    __syncwarp();
    if (t < a.nranks) {
        double *dst = a.mail[t] + rec + (long long)a.rank * (P2P_AR_MAXN + 1);
        for (int i = 0; i < a.n; i++) dst[i] = a.buf[i];
        __threadfence_system();
        p2p_st_flag(reinterpret_cast<unsigned long long *>(dst + P2P_AR_MAXN), a.epoch);
        const double *src = a.mail[a.rank] + rec + (long long)t * (P2P_AR_MAXN + 1);
        if (!p2p_wait(reinterpret_cast<const unsigned long long *>(src + P2P_AR_MAXN), a.epoch, a.err)) ok = 0;
        __threadfence_system();
    }
    __syncwarp();
    if (t < a.n && ok) {
        const double *mine = a.mail[a.rank] + rec;
        double v = __ldcg(mine + t);
        for (int r = 1; r < a.nranks; r++) {
            const double w = __ldcg(mine + (long long)r * (P2P_AR_MAXN + 1) + t);
            v = a.op == RED_SUM ? v + w : (a.op == RED_MAX ? (v > w ? v : w) : (v < w ? v : w));
        }
        a.buf[t] = v;
    }
}
#endif
