// ref_zero_new.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libimpec_ref.so).
// The reference reads members of SimParams it never initialises: the parse of the "levelset" group is commented out
// (src/SimParams.cpp:337-355), yet ls_mod_diff / ls_flagged_diff / sigma0..2 are tested and scaled at :456-466.  Upstream this works
// because a fresh heap page is zero.  To make the UNMODIFIED sources deterministic in a long-lived test process, every object they
// allocate with new comes from zero-filled memory: the replaceable global allocation functions below use calloc.
#include <cstdlib>
#include <new>
void *operator new(std::size_t n) { void *p = std::calloc(n ? n : 1, 1); if (!p) throw std::bad_alloc(); return p; }
void *operator new[](std::size_t n) { void *p = std::calloc(n ? n : 1, 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void *p) noexcept { std::free(p); }
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete(void *p, std::size_t) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }
