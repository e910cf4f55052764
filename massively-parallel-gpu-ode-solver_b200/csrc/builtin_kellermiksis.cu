// Built-in launchers: Keller-Miksis bubble (config C4), without and with dense output.
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

// register cap of these kernels (PerThreadLauncher MINB), measured on B200: 128 registers: 36.2 ms vs 39.1 ms (80, spills) vs 52.3 ms (uncapped, 144) for 2 phases of 2^18 instances
#ifndef MPGOS_MINB_KM
#define MPGOS_MINB_KM 2
#endif

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("kellermiksis", KellerMiksis, double, kAlgRKCK45, false, MPGOS_MINB_KM),
	MPGOS_BUILTIN_ENTRY("kellermiksis", KellerMiksis, double, kAlgRK4, false, MPGOS_MINB_KM),
	MPGOS_BUILTIN_ENTRY("kellermiksis_dense", KellerMiksisDense, double, kAlgRKCK45, true, MPGOS_MINB_KM),
	MPGOS_BUILTIN_ENTRY("kellermiksis", KellerMiksis, float, kAlgRKCK45, false, MPGOS_MINB_KM),
	MPGOS_BUILTIN_ENTRY("kellermiksis_dense", KellerMiksisDense, float, kAlgRKCK45, true, MPGOS_MINB_KM),
	// the same systems with the opt-in transcendentals of include/mpgos_b200/fast_math.cuh
	MPGOS_BUILTIN_ENTRY("kellermiksis_fm", KellerMiksisFm, double, kAlgRKCK45, false, MPGOS_MINB_KM),
	MPGOS_BUILTIN_ENTRY("kellermiksis_dense_fm", KellerMiksisDenseFm, double, kAlgRKCK45, true, MPGOS_MINB_KM),
};
const BuiltinEntry* builtin_entries_kellermiksis(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
