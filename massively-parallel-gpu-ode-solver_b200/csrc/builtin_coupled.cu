// Built-in launchers of the coupled module: the bubble ensemble (config C5) and the ring of Duffing oscillators.
#include "../../include/mpgos_b200/perblock_kernel.cuh"
#include "builtin_coupled_systems.cuh"
#include "builtin_registry.h"

namespace mpgos_b200 {
using namespace builtin;

#define COUPLED_ENTRY_P(NAME, S, P, UPS, CBW, CCI, ALG, D)                                                                    \
	{ NAME, UPS, CBW, CCI, S<P>::UD, S<P>::NC, S<P>::NUP, S<P>::NSP, S<P>::NGP, S<P>::NiGP, S<P>::NUA,                         \
	  S<P>::NiUA, S<P>::NSA, S<P>::NiSA, S<P>::NE, ALG, (int)sizeof(P), D,                                                     \
	  &PerBlockLauncher<S<P>, P, UPS, S<P>::UD, S<P>::NC, CBW, CCI, S<P>::NUP, S<P>::NSP, S<P>::NGP,                           \
	                    S<P>::NiGP, S<P>::NUA, S<P>::NiUA, S<P>::NSA, S<P>::NiSA, S<P>::NE, D, ALG>::launch }
#define COUPLED_ENTRY(NAME, S, UPS, CBW, CCI, ALG, D) COUPLED_ENTRY_P(NAME, S, double, UPS, CBW, CCI, ALG, D)

static const CoupledBuiltinEntry kEntries[] = {
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 8, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 16, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 16, 0, 0, kAlgRK4, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 32, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 64, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 64, 0, 0, kAlgRK4, false),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 64, 0, 0, kAlgRKCK45, true),
	COUPLED_ENTRY("bubble_ensemble", BubbleEnsemble, 128, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("bubble_ensemble_fm", BubbleEnsembleFm, 64, 0, 0, kAlgRKCK45, false),      // opt-in transcendentals (fast_math.cuh)
	COUPLED_ENTRY("bubble_ensemble_fm", BubbleEnsembleFm, 16, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 1, 1, kAlgRKCK45, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 1, 0, kAlgRKCK45, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 0, 1, kAlgRK4, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 14, 0, 0, kAlgRKCK45, true),
	COUPLED_ENTRY("duffing_ring", DuffingRing, 32, 0, 0, kAlgRKCK45, true),      // full matrices, 2 couplings, events: the tensor-core path with everything on
	COUPLED_ENTRY("duffing_ring", DuffingRing, 192, 0, 0, kAlgRKCK45, true),     // two full 192 x 192 matrices (576 KB): more than shared memory holds -> read from global memory
	COUPLED_ENTRY("duffing_ring", DuffingRing, 1500, 1, 1, kAlgRKCK45, true),     // more than 1024 units: two units per thread (768 threads), coupling terms in dynamic shared memory
	COUPLED_ENTRY("duffing_ring", DuffingRing, 2500, 2, 0, kAlgRK4, false),        // three units per thread (864 threads), banded matrices
	COUPLED_ENTRY("parameter_drift", ParameterDrift, 16, 0, 0, kAlgRKCK45, false),      // test system: callbacks write the parameter arrays
	// FP32 instantiations (dense coupling on the tensor cores as 3xTF32 for UPS 32 / 64)
	COUPLED_ENTRY_P("bubble_ensemble", BubbleEnsemble, float, 16, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY_P("bubble_ensemble", BubbleEnsemble, float, 32, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY_P("bubble_ensemble", BubbleEnsemble, float, 64, 0, 0, kAlgRKCK45, false),
	COUPLED_ENTRY_P("bubble_ensemble", BubbleEnsemble, float, 64, 0, 0, kAlgRK4, false),
	COUPLED_ENTRY_P("duffing_ring", DuffingRing, float, 14, 1, 1, kAlgRKCK45, true),
	COUPLED_ENTRY_P("duffing_ring", DuffingRing, float, 14, 0, 0, kAlgRKCK45, true),
};
const CoupledBuiltinEntry* builtin_entries_coupled(int* n) { *n = (int)(sizeof(kEntries) / sizeof(kEntries[0])); return kEntries; }
}
