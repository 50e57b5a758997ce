// Driver.cpp -- `impec2d <config file> [--steps N] [--checkpoint FILE] [--checkpoint-every N] [--restart FILE]`: same CLI as the
// reference's binary (one positional configuration file, src/Driver.cpp:63-67) plus optional overrides: a step limit for smoke
// runs and the binary restart files.
#include "Impec2DSolver.h"
#include <iostream>
#include <cstring>
#include <cstdlib>
int main(int argc, char **argv) {
    if (argc < 2) { std::cerr << "usage: " << argv[0] << " <configuration file> [--steps N]" << std::endl; return 1; }
    Impec2DSolver solver(argv[1], 0, 1);
    for (int a = 2; a + 1 < argc; a += 2)
        if (!strcmp(argv[a], "--steps")) solver.max_steps = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "--checkpoint")) solver.checkpoint_file = argv[a + 1];
        else if (!strcmp(argv[a], "--checkpoint-every")) solver.checkpoint_every = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "--restart")) solver.restart_file = argv[a + 1];
        else { std::cerr << "unknown option " << argv[a] << std::endl; return 1; }
    return solver.Run() ? 0 : 2;
}
