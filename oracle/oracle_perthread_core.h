// TEST INFRASTRUCTURE ONLY -- "parity pinned": this restatement is checked bit for bit against
// oracle/_ref (the reference's own kernel source compiled for the host with the same compiler
// flags) by tests/test_oracle_golden.py (tutorial options) and tests/test_oracle_option_fuzz.py (random solver options), and against the fixtures under tests/golden/
// that were generated from oracle/_ref (tests/golden/make_golden.py).
//
// CPU restatement of the reference's thread-per-system integration phase.  One call of
// integrate_phase() does for one instance what one thread of the reference kernel
//     SourceCodes/SingleSystem_PerThread_Solver.cuh:91-436
// does in one Solve(): load, Initialization, first event values, first dense point, the
// step loop (stepper -> error controller -> event step control -> accept/commit -> callbacks ->
// dense output -> termination), Finalization, store.  Floating-point operations are written in
// the same order as the reference so that, without FMA contraction, results are identical.
#ifndef MPGOS_B200_ORACLE_PERTHREAD_CORE_H
#define MPGOS_B200_ORACLE_PERTHREAD_CORE_H

#include <cmath>
#include <cstdio>

namespace oracle {

enum { ALG_RK4 = 0, ALG_RKCK45 = 1 };

template <class P>
struct PhaseOptions   // Struct_SolverOptions, Interface.cuh:105-115
{
	P   dt_initial, dt_max, dt_min, grow_limit, shrink_limit, dense_min_dt;
	int dense_save_frequency;
	const P*   rtol;      // [SD]
	const P*   atol;      // [SD]
	const P*   ev_tol;    // [NE]
	const int* ev_dir;    // [NE]
	int  ndo;             // dense output capacity per instance
	bool verbose;         // print the reference's device-side warnings
};

// One instance's private copy of everything the reference keeps in registers (Solver.cuh:95-115).
template <class Sys, class P>
struct Instance
{
	enum { SD = Sys::SD, NCP = Sys::NCP, NE = Sys::NE, NA = Sys::NA, NIA = Sys::NIA };
	P   td[2];
	P   x[SD];
	P   cpar[NCP > 0 ? NCP : 1];
	P   acc[NA > 0 ? NA : 1];
	int acci[NIA > 0 ? NIA : 1];
	P   t;
	int dense_index;
	P*  dense_time;    // [ndo]      private, point-major
	P*  dense_state;   // [ndo][SD]
	long long attempted, accepted;
};

// Classic RK4; Steppers.cuh:7-117.  `acc` accumulates k1 + 2 k2 + 2 k3 in place like r_NextState.
template <class Sys, class P>
static inline void stepper_rk4(int tid, int NT, P t, P dt, Instance<Sys, P>& s, P* sp, int* spi, P* xn, bool& finite)
{
	enum { SD = Sys::SD };
	P X[SD], k[SD], T;
	const P half = static_cast<P>(0.5) * dt;
	const P sixth = static_cast<P>(1.0 / 6.0) * dt;

	Sys::ode_function(tid, NT, xn, s.x, t, s.cpar, sp, spi, s.acc, s.acci);
	T = t + half;
	for (int i = 0; i < SD; i++) X[i] = s.x[i] + xn[i] * half;
	Sys::ode_function(tid, NT, k, X, T, s.cpar, sp, spi, s.acc, s.acci);
	for (int i = 0; i < SD; i++) { xn[i] = xn[i] + static_cast<P>(2.0) * k[i]; X[i] = s.x[i] + k[i] * half; }
	Sys::ode_function(tid, NT, k, X, T, s.cpar, sp, spi, s.acc, s.acci);
	T = t + dt;
	for (int i = 0; i < SD; i++) { xn[i] = xn[i] + static_cast<P>(2.0) * k[i]; X[i] = s.x[i] + k[i] * dt; }
	Sys::ode_function(tid, NT, k, X, T, s.cpar, sp, spi, s.acc, s.acci);
	for (int i = 0; i < SD; i++)
	{
		xn[i] = s.x[i] + sixth * (xn[i] + k[i]);
		if (!std::isfinite(xn[i])) finite = false;
	}
}

// Cash-Karp 4(5); Steppers.cuh:122-290.
template <class Sys, class P>
static inline void stepper_rkck45(int tid, int NT, P t, P dt, Instance<Sys, P>& s, P* sp, int* spi, P* xn, P* err, bool& finite)
{
	enum { SD = Sys::SD };
	P X[SD], T, k1[SD], k2[SD], k3[SD], k4[SD], k5[SD], k6[SD];
#define C_(v) static_cast<P>(v)
	Sys::ode_function(tid, NT, k1, s.x, t, s.cpar, sp, spi, s.acc, s.acci);

	T = t + dt * C_(1.0 / 5.0);
	for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (C_(1.0 / 5.0) * k1[i]);
	Sys::ode_function(tid, NT, k2, X, T, s.cpar, sp, spi, s.acc, s.acci);

	T = t + dt * C_(3.0 / 10.0);
	for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (C_(3.0 / 40.0) * k1[i] + C_(9.0 / 40.0) * k2[i]);
	Sys::ode_function(tid, NT, k3, X, T, s.cpar, sp, spi, s.acc, s.acci);

	T = t + dt * C_(3.0 / 5.0);
	for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (C_(3.0 / 10.0) * k1[i] + C_(-9.0 / 10.0) * k2[i] + C_(6.0 / 5.0) * k3[i]);
	Sys::ode_function(tid, NT, k4, X, T, s.cpar, sp, spi, s.acc, s.acci);

	T = t + dt;
	for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (C_(-11.0 / 54.0) * k1[i] + C_(5.0 / 2.0) * k2[i] + C_(-70.0 / 27.0) * k3[i] + C_(35.0 / 27.0) * k4[i]);
	Sys::ode_function(tid, NT, k5, X, T, s.cpar, sp, spi, s.acc, s.acci);

	T = t + dt * C_(7.0 / 8.0);
	for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (C_(1631.0 / 55296.0) * k1[i] + C_(175.0 / 512.0) * k2[i] + C_(575.0 / 13824.0) * k3[i] + C_(44275.0 / 110592.0) * k4[i] + C_(253.0 / 4096.0) * k5[i]);
	Sys::ode_function(tid, NT, k6, X, T, s.cpar, sp, spi, s.acc, s.acci);

	for (int i = 0; i < SD; i++)
	{
		xn[i] = s.x[i] + dt * (C_(37.0 / 378.0) * k1[i] + C_(250.0 / 621.0) * k3[i] + C_(125.0 / 594.0) * k4[i] + C_(512.0 / 1771.0) * k6[i]);
		P e = C_(37.0 / 378.0 - 2825.0 / 27648.0) * k1[i] + C_(250.0 / 621.0 - 18575.0 / 48384.0) * k3[i]
		    + C_(125.0 / 594.0 - 13525.0 / 55296.0) * k4[i] + C_(0.0 - 277.0 / 14336.0) * k5[i] + C_(512.0 / 1771.0 - 1.0 / 4.0) * k6[i];
		// the floor 1e-18 is a double literal in the reference (Steppers.cuh:285): the sum is formed in double
		err[i] = static_cast<P>(dt * std::abs(e) + 1e-18);
		if (!std::isfinite(xn[i]) || !std::isfinite(err[i])) finite = false;
	}
#undef C_
}

// ErrorControllers.cuh:24-86.  Returns through update / terminate / dt_new.
template <int SD, class P>
static inline void controller_rkck45(int tid, P dt, const P* x, const P* xn, const P* err, const PhaseOptions<P>& o,
                                     bool& update, bool& finite, bool& terminate, P& dt_new)
{
	P rel = static_cast<P>(1e30), tol, mult;
	for (int i = 0; i < SD; i++)
	{
		tol = std::fmax(o.rtol[i] * std::fmax(std::abs(xn[i]), std::abs(x[i])), o.atol[i]);
		update = update && (err[i] < tol);
		rel = std::fmin(rel, tol / err[i]);
	}
	if (update) mult = static_cast<P>(0.9) * std::pow(rel, static_cast<P>(1.0 / 5.0));
	else        mult = static_cast<P>(0.9) * std::pow(rel, static_cast<P>(1.0 / 4.0));
	if (!std::isfinite(mult)) finite = false;

	if (!finite)
	{
		mult = o.shrink_limit;
		update = false;
		if (dt < o.dt_min * static_cast<P>(1.01))
		{
			if (o.verbose) std::printf("Error: State is not a finite number even with the minimal step size. (instance %d)\n", tid);
			terminate = true;
		}
	}
	else if (dt < o.dt_min * static_cast<P>(1.01))
	{
		if (o.verbose) std::printf("Warning: Minimum step size reached! (instance %d)\n", tid);
		update = true;
	}
	mult = std::fmin(mult, o.grow_limit);
	mult = std::fmax(mult, o.shrink_limit);
	dt_new = dt * mult;
	dt_new = std::fmin(dt_new, o.dt_max);
	dt_new = std::fmax(dt_new, o.dt_min);
}

// EventHandling.cuh:6-45: secant estimate of the step that lands on the first sign change.
template <int NE, class P>
static inline void event_time_step_control(int tid, bool& update, bool terminate, const P* ev_actual, const P* ev_next,
                                           const PhaseOptions<P>& o, P dt, P& dt_new)
{
	P dt_event = dt;
	bool corrected = false;
	if (update && !terminate)
		for (int i = 0; i < NE; i++)
			if (((ev_actual[i] > o.ev_tol[i]) && (ev_next[i] < -o.ev_tol[i]) && (o.ev_dir[i] <= 0)) ||
			    ((ev_actual[i] < -o.ev_tol[i]) && (ev_next[i] > o.ev_tol[i]) && (o.ev_dir[i] >= 0)))
			{
				dt_event = std::fmin(dt_event, -ev_actual[i] / (ev_next[i] - ev_actual[i]) * dt);
				corrected = true;
			}
	if (corrected)
	{
		if (dt_event < o.dt_min)
		{
			if (o.verbose) std::printf("Warning: Event cannot be detected without reducing the step size below the minimum! (instance %d)\n", tid);
		}
		else { dt_new = dt_event; update = false; }
	}
}

// DenseOutput.cuh:6-37 on the instance-private buffers.
template <class Sys, class P>
static inline void dense_store(Instance<Sys, P>& s, bool store, int& skipped, P& dense_next_time, P dense_min_dt, int ndo)
{
	if (store)
	{
		// The reference stores the first point of a phase unconditionally (Solver.cuh:179-193) and
		// so writes out of bounds once the buffer is full; we keep its bookkeeping (index, skip
		// counter, next time) but drop the out-of-range write itself.
		if (s.dense_index < ndo)
		{
			s.dense_time[s.dense_index] = s.t;
			for (int i = 0; i < Sys::SD; i++) s.dense_state[s.dense_index * Sys::SD + i] = s.x[i];
		}
		s.dense_index++;
		skipped = 0;
		dense_next_time = std::fmin(s.t + dense_min_dt, s.td[1]);
	}
	else skipped++;
}

template <class Sys, class P, int ALG>
static void integrate_phase(int tid, int NT, Instance<Sys, P>& s, P* sp, int* spi, const PhaseOptions<P>& o)
{
	enum { SD = Sys::SD, NE = Sys::NE };
	P xn[SD], err[SD];
	P ev_actual[NE > 0 ? NE : 1], ev_next[NE > 0 ? NE : 1];
	P dt = o.dt_initial, dt_new = o.dt_initial, dense_next_time = 0;
	int skipped = 0, user_terminate = 0;
	bool terminate = false;

	Sys::initialization(tid, NT, s.dense_index, s.t, dt, s.td, s.x, s.cpar, sp, spi, s.acc, s.acci);
	if (NE > 0) Sys::event_function(tid, NT, ev_actual, s.t, dt, s.td, s.x, s.cpar, sp, spi, s.acc, s.acci);
	if (o.ndo > 0) dense_store(s, true, skipped, dense_next_time, o.dense_min_dt, o.ndo);   // Solver.cuh:179-193 (unguarded first point)

	while (!terminate)
	{
		bool update = true, finite = true, end_reached = false;
		dt = dt_new;
		if (dt > (s.td[1] - s.t)) { dt = s.td[1] - s.t; end_reached = true; }

		s.attempted++;
		if (ALG == ALG_RK4)
		{
			stepper_rk4<Sys, P>(tid, NT, s.t, dt, s, sp, spi, xn, finite);
			if (!finite)   // ErrorControllers.cuh:6-20: message, terminate, state still committed
			{
				if (o.verbose) std::printf("Error: State is not a finite number. (instance %d)\n", tid);
				terminate = true;
			}
			dt_new = o.dt_initial;
		}
		else
		{
			stepper_rkck45<Sys, P>(tid, NT, s.t, dt, s, sp, spi, xn, err, finite);
			controller_rkck45<SD, P>(tid, dt, s.x, xn, err, o, update, finite, terminate, dt_new);
		}

		if (NE > 0)
		{
			Sys::event_function(tid, NT, ev_next, s.t + dt, dt, s.td, xn, s.cpar, sp, spi, s.acc, s.acci);
			event_time_step_control<NE, P>(tid, update, terminate, ev_actual, ev_next, o, dt, dt_new);
		}

		if (update)
		{
			s.accepted++;
			s.t += dt;
			for (int i = 0; i < SD; i++) s.x[i] = xn[i];

			Sys::action_after_successful_time_step(tid, NT, user_terminate, s.t, dt, s.td, s.x, s.cpar, sp, spi, s.acc, s.acci);

			if (NE > 0)
			{
				for (int i = 0; i < NE; i++)
					if (((ev_actual[i] > o.ev_tol[i]) && (std::abs(ev_next[i]) < o.ev_tol[i]) && (o.ev_dir[i] <= 0)) ||
					    ((ev_actual[i] < -o.ev_tol[i]) && (std::abs(ev_next[i]) < o.ev_tol[i]) && (o.ev_dir[i] >= 0)))
						Sys::action_after_event_detection(tid, NT, i, user_terminate, s.t, dt, s.td, s.x, s.cpar, sp, spi, s.acc, s.acci);
				Sys::event_function(tid, NT, ev_next, s.t, dt, s.td, s.x, s.cpar, sp, spi, s.acc, s.acci);
				for (int i = 0; i < NE; i++) ev_actual[i] = ev_next[i];
			}

			if (o.ndo > 0)
			{
				// DenseOutput.cuh:41-58
				bool store = (s.dense_index < o.ndo) && (dense_next_time < s.t) && (skipped >= (o.dense_save_frequency - 1));
				if ((s.dense_index < o.ndo) && (end_reached || user_terminate == 1)) store = true;
				dense_store(s, store, skipped, dense_next_time, o.dense_min_dt, o.ndo);
			}

			if (end_reached || user_terminate == 1) terminate = true;
		}
	}

	Sys::finalization(tid, NT, s.dense_index, s.t, dt, s.td, s.x, s.cpar, sp, spi, s.acc, s.acci);
}

} // namespace oracle
#endif
