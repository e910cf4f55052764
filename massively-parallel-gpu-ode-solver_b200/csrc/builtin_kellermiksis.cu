// Built-in launchers: Keller-Miksis bubble (config C4), without and with dense output.
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "builtin_systems.cuh"
#include "builtin_registry.h"

namespace mpgos_b200 {
using namespace builtin;
static const BuiltinEntry kEntries[] = {
	MPGOS_BUILTIN_ENTRY("kellermiksis", KellerMiksis, double, kAlgRKCK45, false),
	MPGOS_BUILTIN_ENTRY("kellermiksis", KellerMiksis, double, kAlgRK4, false),
	MPGOS_BUILTIN_ENTRY("kellermiksis_dense", KellerMiksisDense, double, kAlgRKCK45, true),
	MPGOS_BUILTIN_ENTRY("kellermiksis", KellerMiksis, float, kAlgRKCK45, false),
};
const BuiltinEntry* builtin_entries_kellermiksis(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
