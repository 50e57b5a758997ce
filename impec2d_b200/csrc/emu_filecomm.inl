// emu_filecomm.inl -- IMPEC_EMU builds only (test infrastructure): a communicator between the PROCESSES of an emulated multi-rank run
// that needs no launcher and no Python: every operation writes this rank's payload to a file in IMPEC_EMU_COMM_DIR (temporary name +
// rename, so a file that exists is complete) and polls for the peers' files.  All ranks issue the same sequence of operations (SPMD), so
// a running sequence number names them.  This is what lets the C++ host layer -- Impec2DSolver(cfg, myid, num_procs) as separate
// processes, src/Driver.cpp:58-87 -- be tested on slabs in a container without a GPU (tests/test_host_layer.py); the Python tests keep
// their gloo callbacks (impec_ctx_set_comm_callbacks).  The product library talks NCCL / peer memory and contains none of this.
#include <string>
#include <vector>
#include <cstdio>
#include <chrono>
#include <thread>
#include <unistd.h>

struct EmuFileComm {
    std::string dir; int rank = 0, nranks = 1; long long seq = 0;
    std::vector<std::pair<long long, std::string> > mine;       // files this rank wrote (sequence number, path)

    int gen = 0;                                                // n-th communicator of this process (a re-created context must not meet old files)
    std::string prefix() const { return dir + "/g" + std::to_string(gen) + "_"; }
    std::string name(long long s, int from, int to) const { return prefix() + "op" + std::to_string(s) + "_" + std::to_string(from) + "_" + std::to_string(to); }
    bool put(const std::string &path, const void *data, size_t bytes) {
        const std::string tmp = path + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f) return false;
        const bool ok = bytes == 0 || fwrite(data, 1, bytes, f) == bytes;
        fclose(f);
        if (!ok || rename(tmp.c_str(), path.c_str()) != 0) return false;
        mine.emplace_back(seq, path);
        return true;
    }
    bool get(const std::string &path, void *data, size_t bytes) const {
        const auto t0 = std::chrono::steady_clock::now();
        for (long long spin = 0;; spin++) {
            FILE *f = fopen(path.c_str(), "rb");
            if (f) { const bool ok = bytes == 0 || fread(data, 1, bytes, f) == bytes; fclose(f); return ok; }
            if (spin > 200) std::this_thread::sleep_for(std::chrono::microseconds(200)); else std::this_thread::yield();
            if ((spin & 1023) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return false;     // a peer died
        }
    }
    // after a collective at sequence number s every rank has consumed everything older than s
    void collect(long long s) {
        size_t k = 0;
        for (auto &m : mine) { if (m.first < s) remove(m.second.c_str()); else mine[k++] = m; }
        mine.resize(k);
    }
    // End of the communicator: a rank may only remove files nobody will read any more.  Every rank but 0 signs off with a "bye" file once
    // it has read everything it needs (its own older files go with it); rank 0 waits for all of them and removes what is left.
    void finish() {
        if (rank != 0) {                                        // (what this rank wrote since the last collective may still be unread: rank 0 removes it)
            std::vector<std::pair<long long, std::string> > left; left.swap(mine);
            std::string list;
            for (auto &m : left) list += m.second + "\n";
            put(prefix() + "bye_" + std::to_string(rank), list.data(), list.size());
            return;
        }
        for (int r = 1; r < nranks; r++) {
            const std::string bye = prefix() + "bye_" + std::to_string(r);
            char c;
            if (!get(bye, &c, 0)) continue;                     // (peer died: leave its files)
            FILE *f = fopen(bye.c_str(), "rb");
            if (f) { char line[4096]; while (fgets(line, sizeof(line), f)) { size_t n = strlen(line); if (n && line[n - 1] == '\n') line[n - 1] = 0; if (line[0]) remove(line); } fclose(f); }
            remove(bye.c_str());
        }
        collect(seq + 1);                                       // every peer has signed off: nobody reads this rank's files any more
    }
    static int sendrecv(void *u, const double *slo, double *rlo, const double *shi, double *rhi, long long n) {
        EmuFileComm &c = *(EmuFileComm *)u; const long long s = ++c.seq; const size_t b = (size_t)n * sizeof(double);
        // peers modulo the rank count (a NULL pointer marks a wall; the periodic self test passes all four); "d" = travelling down, "u" = up
        const int lo = (c.rank + c.nranks - 1) % c.nranks, hi = (c.rank + 1) % c.nranks;
        if (slo && !c.put(c.name(s, c.rank, lo) + "d", slo, b)) return 1;
        if (shi && !c.put(c.name(s, c.rank, hi) + "u", shi, b)) return 1;
        if (rlo && !c.get(c.name(s, lo, c.rank) + "u", rlo, b)) return 1;
        if (rhi && !c.get(c.name(s, hi, c.rank) + "d", rhi, b)) return 1;
        return 0;
    }
    static int allgather(void *u, const double *send, double *recv, long long n) {
        EmuFileComm &c = *(EmuFileComm *)u; const long long s = ++c.seq; const size_t b = (size_t)n * sizeof(double);
        if (!c.put(c.name(s, c.rank, -1), send, b)) return 1;
        for (int r = 0; r < c.nranks; r++) {
            if (r == c.rank) { if (recv + (size_t)r * n != send) memcpy(recv + (size_t)r * n, send, b); }
            else if (!c.get(c.name(s, r, -1), recv + (size_t)r * n, b)) return 1;
        }
        c.collect(s);
        return 0;
    }
    static int allreduce(void *u, double *buf, int n, int op) {
        EmuFileComm &c = *(EmuFileComm *)u;
        std::vector<double> all((size_t)n * c.nranks);
        if (allgather(u, buf, all.data(), n)) return 1;
        for (int k = 0; k < n; k++) {                            // rank order on every rank: identical results everywhere
            double a = all[k];
            for (int r = 1; r < c.nranks; r++) { const double v = all[(size_t)r * n + k]; a = op == 0 ? a + v : (op == 1 ? (v > a ? v : a) : (v < a ? v : a)); }
            buf[k] = a;
        }
        return 0;
    }
    static int bcast(void *u, double *buf, long long n, int root) {
        EmuFileComm &c = *(EmuFileComm *)u; const long long s = ++c.seq; const size_t b = (size_t)n * sizeof(double);
        if (c.rank == root) return c.put(c.name(s, root, -2), buf, b) ? 0 : 1;
        return c.get(c.name(s, root, -2), buf, b) ? 0 : 1;
    }
};
