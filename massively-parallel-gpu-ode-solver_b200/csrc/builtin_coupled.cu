// Built-in launchers of the coupled module: the bubble ensemble (config C5) and the ring of Duffing oscillators.
#include "../../include/mpgos_b200/perblock_kernel.cuh"
#include "builtin_coupled_systems.cuh"
#include "builtin_registry.h"

namespace mpgos_b200 {
using namespace builtin;

#define COUPLED_ENTRY(NAME, S, UPS, CBW, CCI, ALG, D)                                                                                      \
	{ NAME, UPS, CBW, CCI, S<double>::UD, S<double>::NC, S<double>::NUP, S<double>::NSP, S<double>::NGP, S<double>::NiGP, S<double>::NUA,     \
	  S<double>::NiUA, S<double>::NSA, S<double>::NiSA, S<double>::NE, ALG, D,                                                             \
	  &PerBlockLauncher<S<double>, double, UPS, S<double>::UD, S<double>::NC, CBW, CCI, S<double>::NUP, S<double>::NSP, S<double>::NGP,    \
	                    S<double>::NiGP, S<double>::NUA, S<double>::NiUA, S<double>::NSA, S<double>::NiSA, S<double>::NE, D, ALG>::launch }

static const CoupledBuiltinEntry kEntries[] = {
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 8, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 16, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 16, 0, 0, kAlgRK4, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 32, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 64, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 64, 0, 0, kAlgRK4, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 64, 0, 0, kAlgRKCK45, true),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 128, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 1, 1, kAlgRKCK45, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 1, 0, kAlgRKCK45, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 0, 1, kAlgRK4, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 0, 0, kAlgRKCK45, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 32, 0, 0, kAlgRKCK45, true),      // full matrices, 2 couplings, events: the tensor-core path with everything on
};
const CoupledBuiltinEntry* builtin_entries_coupled(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
