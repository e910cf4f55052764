// Built-in launchers: pressure relief valve with impact law (config C3).
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRKCK45, false),
	MPGOS_BUILTIN_ENTRY("impact", Impact, double, kAlgRK4, false),
	MPGOS_BUILTIN_ENTRY("impact", Impact, float, kAlgRKCK45, false),
};
const BuiltinEntry* builtin_entries_impact(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
