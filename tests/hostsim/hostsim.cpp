// LOGIC-TEST DOUBLE, build container only (no GPU there).  Compiles the SAME pipeline source as the
// product (phaser_b200/csrc/phz_pipeline.h) against HostSimBackend, which runs every per-thread
// functor serially on the CPU.  It lets `pytest -m "not gpu"` exercise the kernel logic and all
// host code against the oracle before GPU time is spent.  It is never shipped: phaser_b200 only
// ever loads phaser_b200/_phz.so and refuses to run without a CUDA device.
#include "../../phaser_b200/csrc/phz_backend.h"
#define PHZ_BACKEND phz::HostSimBackend
#define PHZ_BACKEND_NAME "hostsim"
#include "../../phaser_b200/csrc/phz_api.inl"
