// Built-in launchers: pressure relief valve with impact law (config C3).
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

// register cap of these kernels (PerThreadLauncher MINB), measured on B200: 80 registers: 52.7 ms vs 55.6 ms (128) vs 63.5 ms (uncapped, 120) for 8 phases of 2^20 instances
#ifndef MPGOS_MINB_IMPACT
#define MPGOS_MINB_IMPACT 3
#endif

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRKCK45, false, MPGOS_MINB_IMPACT),
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRK4, false, MPGOS_MINB_IMPACT),
	MPGOS_BUILTIN_ENTRY("impact", Impact, float, kAlgRKCK45, false, MPGOS_MINB_IMPACT),
};
const BuiltinEntry* builtin_entries_impact(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
