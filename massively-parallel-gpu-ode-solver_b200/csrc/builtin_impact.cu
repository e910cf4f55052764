// Built-in launchers: pressure relief valve with impact law (config C3).
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

// register cap of these kernels (PerThreadLauncher MINB), measured on B200 (8 phases of 2^20 instances, round 2:
// profiles/r2_variant_ab_runs.txt): 96 registers (5 blocks of 128 threads: 4 bytes of spill) 45.4 ms, 80 registers (60 bytes) 46.8 ms,
// 128 registers (none) 47.6 ms
#ifndef MPGOS_MINB_IMPACT
#define MPGOS_MINB_IMPACT 5
#endif
#ifndef MPGOS_MINB_IMPACT_F32
#define MPGOS_MINB_IMPACT_F32 3
#endif

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRKCK45, false, MPGOS_MINB_IMPACT),
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRK4, false, MPGOS_MINB_IMPACT),
	MPGOS_BUILTIN_ENTRY("impact", Impact, float, kAlgRKCK45, false, MPGOS_MINB_IMPACT_F32),
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRKCK45, true, MPGOS_MINB_IMPACT),      // with dense output: lanes drift apart at the impacts (store-efficiency probe)
};
const BuiltinEntry* builtin_entries_impact(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
