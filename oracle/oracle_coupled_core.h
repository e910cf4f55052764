// TEST INFRASTRUCTURE ONLY -- "parity pinned": checked bit for bit against oracle/_ref/libref_coupled_*.so
// (the reference's own coupled kernel source run on the host) by tests/test_oracle_coupled.py and against
// the fixtures under tests/golden/ generated from it.
//
// CPU restatement of the reference's block-per-system integration phase for ONE system: what one
// (logical) block of SourceCodes/CoupledSystmes_PerBlock_Solver.cuh does for one of its systems in one
// Solve() -- :2215-2888 (SingleSystem_SingleBlockLaunch) and, with identical arithmetic per system, the
// three other launch variants (:11, :787, :1508).  Units are visited in ascending order where the
// reference runs them as parallel threads; every cross-unit quantity (accept flag AND, minimum relative
// error, minimum event step) is order independent, and the coupling sum runs over columns in ascending
// order exactly as CoupledSystems_PerBlock_ExplicitRungeKutta_Steppers.cuh:2634-2713.
#ifndef MPGOS_B200_ORACLE_COUPLED_CORE_H
#define MPGOS_B200_ORACLE_COUPLED_CORE_H

#include <cmath>
#include <vector>

namespace oracle_coupled {

enum { ALG_RK4 = 0, ALG_RKCK45 = 1 };

template <class P>
struct Options
{
	P dt_initial, dt_max, dt_min, grow, shrink, dense_min_dt;
	int dense_freq;
	const P* rtol; const P* atol; const P* ev_tol; const int* ev_dir;
	int ndo;
};

template <int N> struct AL1 { enum { v = (N > 0 ? N : 1) }; };

// One system: unit-scope data as [unit][serial] private copies, system scope as plain members.
template <class Sys, class P>
struct System
{
	enum { UPS = Sys::UPS, UD = Sys::UD, NC = Sys::NC, NUP = Sys::NUP, NSP = Sys::NSP, NUA = Sys::NUA, NiUA = Sys::NiUA,
	       NSA = Sys::NSA, NiSA = Sys::NiSA, NE = Sys::NE };
	P   x[UPS][UD];
	P   upar[UPS][AL1<NUP>::v];
	P   uacc[UPS][AL1<NUA>::v];
	int iuacc[UPS][AL1<NiUA>::v];
	P   td[2], t;
	P   spar[AL1<NSP>::v], sacc[AL1<NSA>::v];
	int isacc[AL1<NiSA>::v];
	P   strength[AL1<NC>::v];
	int dense_index;
	P*  dense_time;      // [ndo]
	P*  dense_state;     // [ndo][UPS][UD]
	long long attempted, accepted;
};

// Steppers.cuh:2634-2713 -- one coupling value for (unit, coupling k) from the PACKED matrix.
template <class Sys, class P>
static inline P coupling_value(const P* M, const P cpt[][AL1<Sys::NC>::v], int unit, int k)
{
	enum { UPS = Sys::UPS, CBW = Sys::CBW, CCI = Sys::CCI };
	P v = 0;
	if (CCI == 0 && CBW == 0)
		for (int col = 0; col < UPS; col++) v += M[unit + col * UPS + k * UPS * UPS] * cpt[col][k];
	if (CCI == 0 && CBW > 0)
		for (int d = 0; d < 2 * CBW + 1; d++)
		{
			int s = unit + d - CBW; if (s >= UPS) s -= UPS; if (s < 0) s += UPS;
			v += M[unit + d * UPS + k * UPS * (2 * CBW + 1)] * cpt[s][k];
		}
	if (CCI == 1 && CBW > 0)
		for (int d = 0; d < 2 * CBW + 1; d++)
		{
			int s = unit + d - CBW; if (s >= UPS) s -= UPS; if (s < 0) s += UPS;
			v += M[d + k * (2 * CBW + 1)] * cpt[s][k];
		}
	if (CCI == 1 && CBW == 0)
		for (int d = 0; d < UPS; d++)
		{
			int s = unit + d - CBW; if (s >= UPS) s -= UPS; if (s < 0) s += UPS;
			v += M[d + k * UPS] * cpt[s][k];
		}
	return v;
}

// Right-hand side of every unit including the explicit coupling term (Steppers.cuh: "K1" ... blocks):
// F = f(unit) ; F[CouplingIndex[k]] += CouplingStrength[k]*CPF[k]*CouplingValue.
template <class Sys, class P>
static inline void rhs_all(int sid, System<Sys, P>& s, P F[][Sys::UD], const P X[][Sys::UD], P T, P* gpar, int* igpar, const P* M, const int* cidx)
{
	enum { UPS = Sys::UPS, NC = Sys::NC };
	P cpt[UPS][AL1<NC>::v], cpf[UPS][AL1<NC>::v];
	for (int u = 0; u < UPS; u++)
		Sys::ode_function(sid, u, F[u], const_cast<P*>(X[u]), T, s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc, cpt[u], cpf[u]);
	for (int u = 0; u < UPS; u++)
		for (int k = 0; k < NC; k++)
		{
			P v = coupling_value<Sys, P>(M, cpt, u, k);
			F[u][cidx[k]] += s.strength[k] * cpf[u][k] * v;
		}
}

template <class Sys, class P, int ALG>
static void integrate_phase(int sid, System<Sys, P>& s, P* gpar, int* igpar, const P* M, const int* cidx, const Options<P>& o)
{
	enum { UPS = Sys::UPS, UD = Sys::UD, NE = Sys::NE };
#define C_(v) static_cast<P>(v)
	static thread_local P xn[UPS][UD], err[UPS][UD], X[UPS][UD], k1[UPS][UD], k2[UPS][UD], k3[UPS][UD], k4[UPS][UD], k5[UPS][UD], k6[UPS][UD];
	static thread_local P ev_a[UPS][AL1<NE>::v], ev_n[UPS][AL1<NE>::v];
	P dt = o.dt_initial, dt_new = o.dt_initial, dense_next_time = 0;
	int skipped = 0, udt = 0;
	bool terminated = false;

	for (int u = 0; u < UPS; u++)   // Solver.cuh:2437-2456
		Sys::initialization(sid, u, s.dense_index, s.t, dt, s.td, s.x[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);
	if (NE > 0)
		for (int u = 0; u < UPS; u++)
			Sys::event_function(sid, u, ev_a[u], s.t, dt, s.td, s.x[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);

	auto dense_store = [&](bool store, bool step_updated) {   // DenseOutput.cuh:314-360
		if (store)
		{
			if (s.dense_index < o.ndo)   // the reference stores the first point of a phase unguarded; we drop the out-of-range write
			{
				for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) s.dense_state[(s.dense_index * UPS + u) * UD + i] = s.x[u][i];
				s.dense_time[s.dense_index] = s.t;
			}
			s.dense_index++;
			skipped = 0;
			dense_next_time = std::fmin(s.t + o.dense_min_dt, s.td[1]);
		}
		else if (step_updated) skipped++;
	};
	if (o.ndo > 0) dense_store(true, false);   // s_UpdateDenseOutput = 1 at phase start (Solver.cuh:2318, 2480-2497)

	while (!terminated)
	{
		bool update = true, finite = true, end_reached = false;
		dt = dt_new;
		if (dt > (s.td[1] - s.t)) { dt = s.td[1] - s.t; end_reached = true; }
		s.attempted++;

		if (ALG == ALG_RK4)
		{
			// Steppers.cuh:2015-2227; xn accumulates k1 + 2 k2 + 2 k3
			rhs_all<Sys, P>(sid, s, xn, s.x, s.t, gpar, igpar, M, cidx);
			P T = s.t + C_(0.5) * dt;
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) X[u][i] = s.x[u][i] + xn[u][i] * C_(0.5) * dt;
			rhs_all<Sys, P>(sid, s, k1, X, T, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) { xn[u][i] += 2 * k1[u][i]; X[u][i] = s.x[u][i] + k1[u][i] * C_(0.5) * dt; }
			rhs_all<Sys, P>(sid, s, k1, X, T, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) { xn[u][i] += C_(2.0) * k1[u][i]; X[u][i] = s.x[u][i] + k1[u][i] * dt; }
			T = s.t + dt;
			rhs_all<Sys, P>(sid, s, k1, X, T, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++)
			{
				xn[u][i] = s.x[u][i] + dt * (xn[u][i] + k1[u][i]) * C_(1.0 / 6.0);
				if (!std::isfinite(xn[u][i])) finite = false;
			}
			// ErrorControllers.cuh:421-441: message, terminate, UpdateStep = 0 (the state is NOT committed here, unlike per-thread)
			if (!finite) { terminated = true; update = false; }
			dt_new = o.dt_initial;
		}
		else
		{
			// Steppers.cuh:2228-2544
			rhs_all<Sys, P>(sid, s, k1, s.x, s.t, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) X[u][i] = s.x[u][i] + (k1[u][i] * C_(1.0 / 5.0)) * dt;
			rhs_all<Sys, P>(sid, s, k2, X, s.t + C_(1.0 / 5.0) * dt, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) X[u][i] = s.x[u][i] + (k1[u][i] * C_(3.0 / 40.0) + k2[u][i] * C_(9.0 / 40.0)) * dt;
			rhs_all<Sys, P>(sid, s, k3, X, s.t + C_(3.0 / 10.0) * dt, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) X[u][i] = s.x[u][i] + (k1[u][i] * C_(3.0 / 10.0) + k2[u][i] * C_(-9.0 / 10.0) + k3[u][i] * C_(6.0 / 5.0)) * dt;
			rhs_all<Sys, P>(sid, s, k4, X, s.t + C_(3.0 / 5.0) * dt, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) X[u][i] = s.x[u][i] + (k1[u][i] * C_(-11.0 / 54.0) + k2[u][i] * C_(5.0 / 2.0) + k3[u][i] * C_(-70.0 / 27.0) + k4[u][i] * C_(35.0 / 27.0)) * dt;
			rhs_all<Sys, P>(sid, s, k5, X, s.t + dt, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) X[u][i] = s.x[u][i] + (k1[u][i] * C_(1631.0 / 55296.0) + k2[u][i] * C_(175.0 / 512.0) + k3[u][i] * C_(575.0 / 13824.0) + k4[u][i] * C_(44275.0 / 110592.0) + k5[u][i] * C_(253.0 / 4096.0)) * dt;
			rhs_all<Sys, P>(sid, s, k6, X, s.t + C_(7.0 / 8.0) * dt, gpar, igpar, M, cidx);
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++)
			{
				xn[u][i] = s.x[u][i] + (k1[u][i] * C_(37.0 / 378.0) + k3[u][i] * C_(250.0 / 621.0) + k4[u][i] * C_(125.0 / 594.0) + k6[u][i] * C_(512.0 / 1771.0)) * dt;
				P e = k1[u][i] * C_(37.0 / 378.0 - 2825.0 / 27648.0) + k3[u][i] * C_(250.0 / 621.0 - 18575.0 / 48384.0) + k4[u][i] * C_(125.0 / 594.0 - 13525.0 / 55296.0)
				    + k5[u][i] * C_(0.0 - 277.0 / 14336.0) + k6[u][i] * C_(512.0 / 1771.0 - 1.0 / 4.0);
				err[u][i] = dt * std::abs(e) + C_(1e-18);
				if (!std::isfinite(xn[u][i]) || !std::isfinite(err[u][i])) finite = false;
			}
			// ErrorControllers.cuh:454-542 -- per-system reduction, safety 0.8
			P rel = C_(1e30);
			for (int u = 0; u < UPS; u++)
				for (int i = 0; i < UD; i++)
				{
					P tol = std::fmax(o.rtol[i] * std::fmax(std::abs(xn[u][i]), std::abs(s.x[u][i])), o.atol[i]);
					update = update && (err[u][i] < tol);
					rel = std::fmin(rel, tol / err[u][i]);
				}
			P mult = update ? C_(0.8) * std::pow(rel, C_(1.0 / 5.0)) : C_(0.8) * std::pow(rel, C_(1.0 / 4.0));
			if (!std::isfinite(mult)) finite = false;
			if (!finite)
			{
				mult = o.shrink;
				update = false;
				if (dt < o.dt_min * C_(1.01)) terminated = true;
			}
			else if (dt < o.dt_min * C_(1.01)) update = true;
			mult = std::fmin(mult, o.grow);
			mult = std::fmax(mult, o.shrink);
			dt_new = dt * mult;
			dt_new = std::fmin(dt_new, o.dt_max);
			dt_new = std::fmax(dt_new, o.dt_min);
		}

		if (NE > 0)   // Solver.cuh:2630-2667, EventHandling.cuh:216-268
		{
			if (!terminated)
				for (int u = 0; u < UPS; u++)
					Sys::event_function(sid, u, ev_n[u], s.t + dt, dt_new, s.td, xn[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);
			P dt_event = dt;
			bool corrected = false;
			if (update && !terminated)
				for (int u = 0; u < UPS; u++)
					for (int i = 0; i < NE; i++)
						if (((ev_a[u][i] > o.ev_tol[i]) && (ev_n[u][i] < -o.ev_tol[i]) && (o.ev_dir[i] <= 0)) ||
						    ((ev_a[u][i] < -o.ev_tol[i]) && (ev_n[u][i] > o.ev_tol[i]) && (o.ev_dir[i] >= 0)))
						{
							dt_event = std::fmin(dt_event, -ev_a[u][i] / (ev_n[u][i] - ev_a[u][i]) * dt);
							corrected = true;
						}
			if (corrected && !(dt_event < o.dt_min)) { dt_new = dt_event; update = false; }
		}

		if (update)   // Solver.cuh:2670-2806
		{
			s.accepted++;
			s.t += dt;
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) s.x[u][i] = xn[u][i];
			for (int u = 0; u < UPS; u++)
				Sys::action_after_successful_time_step(sid, u, udt, s.t, dt, s.td, s.x[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);
			if (NE > 0)
			{
				for (int u = 0; u < UPS; u++)
					for (int i = 0; i < NE; i++)
						if (((ev_a[u][i] > o.ev_tol[i]) && (std::abs(ev_n[u][i]) < o.ev_tol[i]) && (o.ev_dir[i] <= 0)) ||
						    ((ev_a[u][i] < -o.ev_tol[i]) && (std::abs(ev_n[u][i]) < o.ev_tol[i]) && (o.ev_dir[i] >= 0)))
							Sys::action_after_event_detection(sid, u, i, udt, s.t, dt, s.td, s.x[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);
				for (int u = 0; u < UPS; u++)
				{
					Sys::event_function(sid, u, ev_n[u], s.t, dt_new, s.td, s.x[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);
					for (int i = 0; i < NE; i++) ev_a[u][i] = ev_n[u][i];
				}
			}
		}
		if (o.ndo > 0)   // DenseOutput.cuh:362-395 then :314-360
		{
			bool store = false;
			if (update)
			{
				store = (s.dense_index < o.ndo) && (dense_next_time < s.t) && (skipped >= (o.dense_freq - 1));
				if ((s.dense_index < o.ndo) && (end_reached || udt == 1)) store = true;
			}
			dense_store(store, update);
		}
		if (update && (end_reached || udt == 1)) terminated = true;
	}
	for (int u = 0; u < UPS; u++)
		Sys::finalization(sid, u, s.dense_index, s.t, dt, s.td, s.x[u], s.upar[u], s.spar, gpar, igpar, s.uacc[u], s.iuacc[u], s.sacc, s.isacc);
#undef C_
}

} // namespace oracle_coupled
#endif
