// Built-in launchers: Duffing with the full T1 feature set (config C1) and its events-free
// variant "poincare" (T3; strict fixed-step gate).
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

// register cap of these kernels (PerThreadLauncher MINB), measured on B200: same class of small polynomial right-hand sides as Lorenz
#ifndef MPGOS_MINB_DUFFING
#define MPGOS_MINB_DUFFING 3
#endif

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("poincare", Poincare, double, kAlgRK4, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("poincare", Poincare, double, kAlgRKCK45, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("poincare", Poincare, float, kAlgRK4, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("poincare", Poincare, float, kAlgRKCK45, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("duffing", Duffing, double, kAlgRK4, true, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("duffing", Duffing, double, kAlgRKCK45, true, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("duffing", Duffing, double, kAlgRK4, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("duffing", Duffing, double, kAlgRKCK45, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("duffing", Duffing, float, kAlgRKCK45, true, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("duffing", Duffing, float, kAlgRK4, true, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("tidprobe", TidProbe, double, kAlgRKCK45, false, MPGOS_MINB_DUFFING),
	MPGOS_BUILTIN_ENTRY("tidprobe", TidProbe, double, kAlgRK4, false, MPGOS_MINB_DUFFING),
};
const BuiltinEntry* builtin_entries_duffing(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
