// Driver.cpp -- `impec2d <config file> [--steps N] [--checkpoint FILE] [--checkpoint-every N] [--restart FILE]`: same CLI as the
// reference's binary (one positional configuration file, src/Driver.cpp:63-67) plus optional overrides: a step limit for smoke
// runs and the binary restart files.
#include "Impec2DSolver.h"
#include <iostream>
#include <cstring>
#include <cstdlib>
#include <initializer_list>
int main(int argc, char **argv) {
    if (argc < 2) { std::cerr << "usage: " << argv[0] << " <configuration file> [--steps N]" << std::endl; return 1; }
    // One process per GPU.  The reference takes rank and size from MPI_Comm_rank / MPI_Comm_size (src/Driver.cpp:58-62); without an MPI
    // launcher they come from the environment of whatever starts the processes: IMPEC_RANK / IMPEC_NRANKS, torchrun's RANK / WORLD_SIZE
    // (`python -m torch.distributed.run --no-python --nproc-per-node N impec2d file.cfg`), Open MPI's or PMI's variables.
    auto env_int = [](std::initializer_list<const char *> names, int dflt) { for (const char *n : names) if (const char *e = getenv(n)) return atoi(e); return dflt; };
    const int myid = env_int({"IMPEC_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK"}, 0);
    const int num_procs = env_int({"IMPEC_NRANKS", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE"}, 1);
    if (myid < 0 || myid >= num_procs) { std::cerr << "bad rank " << myid << " of " << num_procs << std::endl; return 1; }
    Impec2DSolver solver(argv[1], myid, num_procs);
    for (int a = 2; a + 1 < argc; a += 2)
        if (!strcmp(argv[a], "--steps")) solver.max_steps = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "--checkpoint")) solver.checkpoint_file = argv[a + 1];
        else if (!strcmp(argv[a], "--checkpoint-every")) solver.checkpoint_every = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "--restart")) solver.restart_file = argv[a + 1];
        else { std::cerr << "unknown option " << argv[a] << std::endl; return 1; }
    return solver.Run() ? 0 : 2;
}
