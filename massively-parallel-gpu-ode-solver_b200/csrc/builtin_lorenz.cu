// Built-in launchers: Lorenz (config C2), RK4 and RKCK45, FP64 and FP32.
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

// register cap of these kernels (PerThreadLauncher MINB), measured on B200 (2^20 instances, round 2: profiles/r2_variant_ab_runs.txt):
// FP64 96 registers (5 blocks of 128 threads: no spill) 4.727 ms, 80 registers (12 bytes of spill) 4.751 ms, 128 registers 4.729 ms;
// FP32 fits 80 registers
#ifndef MPGOS_MINB_LORENZ
#define MPGOS_MINB_LORENZ 5
#endif
#ifndef MPGOS_MINB_LORENZ_F32
#define MPGOS_MINB_LORENZ_F32 3
#endif

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, double, kAlgRK4, false, MPGOS_MINB_LORENZ),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, double, kAlgRKCK45, false, MPGOS_MINB_LORENZ),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, float, kAlgRK4, false, MPGOS_MINB_LORENZ_F32),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, float, kAlgRKCK45, false, MPGOS_MINB_LORENZ_F32),
};
const BuiltinEntry* builtin_entries_lorenz(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
