// builtin_registry.h -- table of the launchers compiled into the library, one translation unit
// per system family so that they build in parallel.
#ifndef MPGOS_B200_BUILTIN_REGISTRY_H
#define MPGOS_B200_BUILTIN_REGISTRY_H

#include "../../include/mpgos_b200.h"
#include "../../include/mpgos_b200_coupled.h"

namespace mpgos_b200 {

struct BuiltinEntry
{
	const char* name;
	int SD, NCP, NSP, NISP, NE, NA, NIA;
	int Algorithm, PrecisionBytes;
	bool Dense;
	mpgos_perthread_launcher launcher;
};

const BuiltinEntry* builtin_entries_lorenz(int* n);
const BuiltinEntry* builtin_entries_duffing(int* n);        // "poincare" and "duffing"
const BuiltinEntry* builtin_entries_impact(int* n);
const BuiltinEntry* builtin_entries_kellermiksis(int* n);   // "kellermiksis" and "kellermiksis_dense"

struct CoupledBuiltinEntry
{
	const char* name;
	int UPS, CBW, CCI, UD, NC, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE;
	int Algorithm, PrecisionBytes;
	bool Dense;
	mpgos_coupled_launcher launcher;
};
const CoupledBuiltinEntry* builtin_entries_coupled(int* n);

} // namespace mpgos_b200

// Entry for system template S (builtin_systems.cuh), precision P, algorithm ALG, dense flag D, register cap MINB
// (resident 256-thread blocks per SM, see PerThreadLauncher).
#define MPGOS_BUILTIN_ENTRY(NAME, S, P, ALG, D, MINB)                                                                 \
	{ NAME, S<P>::SD, S<P>::NCP, S<P>::NSP, S<P>::NISP, S<P>::NE, S<P>::NA, S<P>::NIA, ALG, (int)sizeof(P), D,        \
	  &mpgos_b200::PerThreadLauncher<S<P>, P, S<P>::SD, S<P>::NCP, S<P>::NSP, S<P>::NISP, S<P>::NE, S<P>::NA, S<P>::NIA, D, ALG, MINB>::launch }

#endif
