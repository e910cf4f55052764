// Built-in launchers: Lorenz (config C2), RK4 and RKCK45, FP64 and FP32.
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

// register cap of these kernels (PerThreadLauncher MINB), measured on B200: 80 registers: 5.39 ms vs 5.52 ms (128) for 2^20 instances
#ifndef MPGOS_MINB_LORENZ
#define MPGOS_MINB_LORENZ 3
#endif

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, double, kAlgRK4, false, MPGOS_MINB_LORENZ),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, double, kAlgRKCK45, false, MPGOS_MINB_LORENZ),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, float, kAlgRK4, false, MPGOS_MINB_LORENZ),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, float, kAlgRKCK45, false, MPGOS_MINB_LORENZ),
};
const BuiltinEntry* builtin_entries_lorenz(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
