// Built-in launchers: Lorenz (config C2), RK4 and RKCK45, FP64 and FP32.
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, double, kAlgRK4, false),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, double, kAlgRKCK45, false),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, float, kAlgRK4, false),
	MPGOS_BUILTIN_ENTRY("lorenz", Lorenz, float, kAlgRKCK45, false),
};
const BuiltinEntry* builtin_entries_lorenz(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
