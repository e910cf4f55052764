// runtime.cu -- implementation of the C-ABI in include/mpgos_b200.h.
//
// Plays the role of the host half of the reference's ProblemSolver object
// (SourceCodes/SingleSystem_PerThread_Interface.cuh:309-1312): device selection, one stream and
// one event per solver object, pinned host mirrors, device buffers in the same structure-of-arrays
// layout, option bookkeeping, asynchronous copies, launch, synchronisation.  The integration
// kernel itself is not here: Solve() receives a launcher generated next to the user's callbacks
// (include/mpgos_b200/perthread_kernel.cuh) or one of the built-in systems (builtin_*.cu).
//
// There is no CPU path: every entry point that needs the device fails with MPGOS_ERR_CUDA when
// CUDA is unavailable.
#include "../../include/mpgos_b200.h"
#include "../../include/mpgos_b200/fast_math.cuh"
#include "builtin_registry.h"

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
	return code;
}

} // namespace
namespace mpgos_b200 {
// shared with runtime_coupled.cu
int set_error(int code, const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
	return code;
}
}
namespace {

#define CUDA_TRY(call)                                                                                   \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess)                                                                           \
			return fail(MPGOS_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

// One host mirror + device buffer pair.
struct Mirror
{
	void*  host = nullptr;
	void*  dev = nullptr;
	size_t elements = 0;
	size_t bytes = 0;
};

// sum of two u64 arrays, one block-level reduction per 1024 elements, finished with atomics
__global__ void reduce_step_counters(const unsigned long long* __restrict__ attempted, const unsigned long long* __restrict__ accepted,
                                     int n, unsigned long long* __restrict__ out)
{
	unsigned long long a = 0, b = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a += attempted[i]; b += accepted[i]; }
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
	if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], a); atomicAdd(&out[1], b); }
}

// Peak probes (the roofline denominators).  Independent FMA chains held in registers: 16 chains x 16 unrolled trips = 256
// FMAs per loop trip, so that the loop's own counter/branch instructions (which cost issue slots the FP64 pipe cannot hide:
// an FP64 instruction takes two issue cycles per sub-partition and other instructions do not overlap with it) are < 1 % of
// the stream.  Round 1's probe had 16 FMAs per trip and read 33.7 TFLOP/s, 9 % below the issue rate the cycle-accurate
// micro-benchmark proves (one DFMA per 2.00 cycles per sub-partition = 37.2 TFLOP/s at 1965 MHz, tools/microbench_fp64.cu).
template <class P>
__global__ void __launch_bounds__(256) fma_peak_kernel(P* out, int trips, P a, P b)
{
	P x[16];
	#pragma unroll
	for (int i = 0; i < 16; i++) x[i] = (P)(threadIdx.x + i);
	for (int it = 0; it < trips; it++)
	{
		#pragma unroll
		for (int u = 0; u < 16; u++)
		{
			#pragma unroll
			for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
		}
	}
	P s = 0;
	#pragma unroll
	for (int i = 0; i < 16; i++) s += x[i];
	if (s == (P)123.456) out[0] = s;   // never true; keeps the chains alive
}

// FP64 tensor-core instruction (mma.sync.m8n8k4.f64, DMMA): 256 FMAs per warp instruction, 4 accumulator pairs in flight
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int trips)
{
	double c0[4], c1[4];
	const double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
	#pragma unroll
	for (int i = 0; i < 4; i++) { c0[i] = i; c1[i] = -i; }
	for (int it = 0; it < trips; it++)
	{
		#pragma unroll
		for (int u = 0; u < 16; u++)
		{
			#pragma unroll
			for (int i = 0; i < 4; i++)
				asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
		}
	}
	double s = 0;
	#pragma unroll
	for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
	if (s == 123.456) out[0] = s;
}

// Times `launch(trips)` (best of 5 after one warm-up launch, CUDA events; each launch lasts ~10 ms so that the clocks are at
// their load level and launch overhead is < 0.1 %) and returns flops / second / 1e12.
template <class Launch>
int time_probe(Launch launch, double flops_per_trip_total, int trips, double* tflops)
{
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));  CUDA_TRY(cudaEventCreate(&e1));
	double best = 0;
	for (int rep = 0; rep < 6; rep++)
	{
		CUDA_TRY(cudaEventRecord(e0));
		launch(trips);
		CUDA_TRY(cudaEventRecord(e1));
		CUDA_TRY(cudaEventSynchronize(e1));
		CUDA_TRY(cudaGetLastError());
		float ms = 0;
		CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
		const double tf = flops_per_trip_total * trips / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);  cudaEventDestroy(e1);
	*tflops = best;
	return MPGOS_OK;
}

template <class P>
int probe_fma_peak(double* tflops)
{
	P* d = nullptr;
	CUDA_TRY(cudaMalloc(&d, sizeof(P)));
	cudaDeviceProp prop;  int dev = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
	const int blocks = prop.multiProcessorCount * 8, threads = 256;
	const int trips = sizeof(P) == 8 ? 1536 : 3072;          // ~10 ms per launch
	const int rc = time_probe([&](int n) { fma_peak_kernel<P><<<blocks, threads>>>(d, n, (P)0.999999, (P)1e-7); },
	                          2.0 * 256.0 * (double)blocks * threads, trips, tflops);
	cudaFree(d);
	return rc;
}

int probe_dmma_peak(double* tflops)
{
	double* d = nullptr;
	CUDA_TRY(cudaMalloc(&d, sizeof(double)));
	cudaDeviceProp prop;  int dev = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
	const int blocks = prop.multiProcessorCount * 8, threads = 256;
	// one DMMA = 8 x 8 x 4 FMAs per warp; 64 DMMAs per trip and warp
	const int rc = time_probe([&](int n) { dmma_peak_kernel<<<blocks, threads>>>(d, n); },
	                          2.0 * 256.0 * 64.0 * (double)blocks * (threads / 32), 768, tflops);
	cudaFree(d);
	return rc;
}

} // namespace

struct mpgos_perthread_solver
{
	mpgos_perthread_desc desc;
	int          sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t  event = nullptr;
	cudaEvent_t  t0 = nullptr, t1 = nullptr;
	Mirror       var[13];   // indexed by mpgos_variable (All and DenseOutput unused)
	unsigned long long* d_attempted = nullptr;
	unsigned long long* d_accepted = nullptr;
	unsigned long long* d_totals = nullptr;   // [2]
	unsigned int*       d_queue = nullptr;
	unsigned int*       d_last_steps = nullptr;   // [NT] steps of the previous phase (cost history of the dispatch order)
	int*                d_order = nullptr;        // [NT/32 + 1]
	unsigned int*       d_order_scratch = nullptr;   // [NT/32 + 1 + 768]
	bool                history_valid = false;    // d_last_steps describes the previous phase of the instances now on the device
	void*               snapshot[13] = {};        // device-side checkpoint, allocated on first save
	// options (Interface.cuh:475-496 defaults)
	double InitialTimeStep = 1e-2, MaximumTimeStep = 1.0e6, MinimumTimeStep = 1.0e-12;
	double TimeStepGrowLimit = 5.0, TimeStepShrinkLimit = 0.1, DenseOutputMinimumTimeStep = 0.0;
	int    DenseOutputSaveFrequency = 1;
	int    ThreadsPerBlock = 128;
	int    ActiveNumberOfThreads = 0;
	int    PreferSharedMemory = 1;
	int    LaneRefill = 0;      // measured on B200: the direct mapping wins once the register cap gives >= 5 warps per sub-partition (profiles/)
	int    DeviceMessages = 0;
	int    DispatchOrder = 1;
	std::vector<double> rtol, atol, evtol;
	std::vector<int>    evdir;
	// shared parameters as of the last host-to-device copy: small sets reach the kernel by value (perthread_kernel.cuh)
	std::vector<double> shared_values;
	std::vector<int>    integer_shared_values;
	// multi-device solver (mpgos_perthread_create_sharded): the parent owns the host mirrors (global layout) and no device
	// memory; each part is an ordinary solver object on its device without host mirrors
	std::vector<mpgos_perthread_solver*> shards;
	int  tile = 0;
	bool has_host_mirrors = true;
};

namespace {

// ---- tile-cyclic shard map (mpgos_b200.h): tile t of `tile` consecutive instances belongs to shard t % count ----
int shard_instances(int NT, int tile, int count, int index)
{
	const int ntiles = (NT + tile - 1) / tile;
	if (index >= ntiles) return 0;
	const int mine = (ntiles - index + count - 1) / count;
	int n = mine * tile;
	if ((ntiles - 1) % count == index) n -= ntiles * tile - NT;      // the sweep's last tile may be short
	return n;
}
long long shard_global_id(int tile, int count, int index, int local) { return ((long long)(local / tile) * count + index) * tile + local % tile; }
// local instances of the shard whose global id is below `limit`
int shard_instances_below(int NT, int tile, int count, int index, int limit)
{
	return shard_instances(limit < NT ? (limit > 0 ? limit : 0) : NT, tile, count, index);
}

// elements per instance of a per-instance variable (0: not per instance)
int serials_of(const mpgos_perthread_desc& d, int v)
{
	switch (v)
	{
		case MPGOS_VAR_TimeDomain: return 2;            case MPGOS_VAR_ActualState: return d.SD;        case MPGOS_VAR_ActualTime: return 1;
		case MPGOS_VAR_ControlParameters: return d.NCP; case MPGOS_VAR_Accessories: return d.NA;        case MPGOS_VAR_IntegerAccessories: return d.NIA;
		case MPGOS_VAR_DenseIndex: return 1;            case MPGOS_VAR_DenseTime: return d.NDO;         case MPGOS_VAR_DenseState: return d.NDO * d.SD;
		default: return 0;
	}
}
size_t element_size_of(const mpgos_perthread_desc& d, int v)
{
	return (v == MPGOS_VAR_IntegerSharedParameters || v == MPGOS_VAR_IntegerAccessories || v == MPGOS_VAR_DenseIndex) ? 4 : (size_t)d.PrecisionBytes;
}

// Strided copy between a host array in the sweep's layout (element (g, serial) at [g + serial * NT]) and one part's device
// buffer (element (i, serial) at [i + serial * n]): one 2-D copy for the whole variable when the sweep is a whole number of
// rounds of tiles, else one per serial plus the short last tile.  Asynchronous on the part's stream.
int copy_shard(mpgos_perthread_solver* parent, int index, int v, unsigned char* host, bool to_device)
{
	mpgos_perthread_solver* c = parent->shards[index];
	const mpgos_perthread_desc& pd = parent->desc;
	const int count = (int)parent->shards.size(), tile = parent->tile, NT = pd.NT, n = c->desc.NT;
	const int serials = serials_of(pd, v);
	const size_t es = element_size_of(pd, v);
	unsigned char* dev = (unsigned char*)c->var[v].dev;
	if (n == 0 || serials == 0 || !dev) return MPGOS_OK;
	CUDA_TRY(cudaSetDevice(c->desc.Device));
	auto copy2d = [&](size_t dev_off, size_t host_off, size_t width, size_t rows) -> cudaError_t
	{
		if (rows == 0 || width == 0) return cudaSuccess;
		return to_device ? cudaMemcpy2DAsync(dev + dev_off, (size_t)tile * es, host + host_off, (size_t)tile * count * es, width, rows, cudaMemcpyHostToDevice, c->stream)
		                 : cudaMemcpy2DAsync(host + host_off, (size_t)tile * count * es, dev + dev_off, (size_t)tile * es, width, rows, cudaMemcpyDeviceToHost, c->stream);
	};
	const size_t first = (size_t)index * tile * es;       // host offset of the part's first tile
	if (NT % (tile * count) == 0)
	{
		// rows (serial, local tile) have a uniform pitch on both sides
		CUDA_TRY(copy2d(0, first, (size_t)tile * es, (size_t)serials * (n / tile)));
		return MPGOS_OK;
	}
	const int full = n / tile, rest = n % tile;
	for (int k = 0; k < serials; k++)
	{
		CUDA_TRY(copy2d((size_t)k * n * es, (size_t)k * NT * es + first, (size_t)tile * es, (size_t)full));
		if (rest) CUDA_TRY(copy2d(((size_t)k * n + (size_t)full * tile) * es, (size_t)k * NT * es + first + (size_t)full * tile * count * es, (size_t)rest * es, 1));
	}
	return MPGOS_OK;
}

} // namespace


extern "C" {

const char* mpgos_last_error(void) { return g_error; }
int mpgos_abi_version(void) { return MPGOS_B200_ABI_VERSION; }

int mpgos_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

static int alloc_mirror(Mirror& m, size_t elements, size_t elem_size, bool host, bool device)
{
	m.elements = elements;
	m.bytes = elements * elem_size;
	if (m.bytes == 0) return MPGOS_OK;
	if (host)
	{
		CUDA_TRY(cudaMallocHost(&m.host, m.bytes));
		std::memset(m.host, 0, m.bytes);
	}
	if (device)
	{
		CUDA_TRY(cudaMalloc(&m.dev, m.bytes));
		CUDA_TRY(cudaMemset(m.dev, 0, m.bytes));
	}
	return MPGOS_OK;
}

// host_mirrors = false: a part of a multi-device solver (its parent owns the host side)
static int create_impl(const mpgos_perthread_desc* desc, bool host_mirrors, mpgos_perthread_solver** out)
{
	if (!desc || !out) return fail(MPGOS_ERR_ARGUMENT, "null argument");
	*out = nullptr;
	const mpgos_perthread_desc& d = *desc;
	if (d.NT <= 0 || d.SD <= 0 || d.NCP < 0 || d.NSP < 0 || d.NISP < 0 || d.NE < 0 || d.NA < 0 || d.NIA < 0 || d.NDO < 0)
		return fail(MPGOS_ERR_ARGUMENT, "invalid sizes in descriptor");
	if (d.Algorithm != MPGOS_RK4 && d.Algorithm != MPGOS_RKCK45) return fail(MPGOS_ERR_ARGUMENT, "unknown algorithm %d", d.Algorithm);
	if (d.PrecisionBytes != 8 && d.PrecisionBytes != 4) return fail(MPGOS_ERR_ARGUMENT, "precision must be 8 or 4 bytes");

	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	if (d.Device < 0 || d.Device >= ndev) return fail(MPGOS_ERR_ARGUMENT, "device %d not present (%d devices)", d.Device, ndev);
	CUDA_TRY(cudaSetDevice(d.Device));

	// memory check before allocating anything (Interface.cuh:368-396)
	const size_t P = (size_t)d.PrecisionBytes, NT = (size_t)d.NT;
	const size_t need = P * NT * (2 + d.SD + 1 + d.NCP + d.NA + (size_t)d.NDO * (1 + d.SD)) + P * d.NSP
	                  + 4 * (d.NISP + NT * d.NIA + NT) + 16 * NT + 64;
	size_t free_b = 0, total_b = 0;
	CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
	if (need >= free_b) return fail(MPGOS_ERR_MEMORY, "solver needs %zu bytes of device memory, %zu free", need, free_b);

	mpgos_perthread_solver* s = new (std::nothrow) mpgos_perthread_solver();
	if (!s) return fail(MPGOS_ERR_MEMORY, "host allocation failed");
	s->desc = d;
	s->ActiveNumberOfThreads = d.NT;
	s->rtol.assign(d.SD, 1e-8);  s->atol.assign(d.SD, 1e-8);
	s->evtol.assign(d.NE > 0 ? d.NE : 1, 1e-6);  s->evdir.assign(d.NE > 0 ? d.NE : 1, 0);
	s->shared_values.assign(d.NSP > 0 ? d.NSP : 1, 0.0);  s->integer_shared_values.assign(d.NISP > 0 ? d.NISP : 1, 0);
	s->has_host_mirrors = host_mirrors;

	int rc = MPGOS_OK;
	auto A = [&](int v, size_t elements, size_t elem) { if (rc == MPGOS_OK) rc = alloc_mirror(s->var[v], elements, elem, host_mirrors, true); };
	A(MPGOS_VAR_TimeDomain, 2 * NT, P);
	A(MPGOS_VAR_ActualState, d.SD * NT, P);
	A(MPGOS_VAR_ActualTime, NT, P);
	A(MPGOS_VAR_ControlParameters, d.NCP * NT, P);
	A(MPGOS_VAR_SharedParameters, d.NSP, P);
	A(MPGOS_VAR_IntegerSharedParameters, d.NISP, 4);
	A(MPGOS_VAR_Accessories, d.NA * NT, P);
	A(MPGOS_VAR_IntegerAccessories, d.NIA * NT, 4);
	A(MPGOS_VAR_DenseIndex, NT, 4);
	A(MPGOS_VAR_DenseTime, (size_t)d.NDO * NT, P);
	A(MPGOS_VAR_DenseState, (size_t)d.NDO * d.SD * NT, P);
	cudaDeviceProp prop;
	if (rc == MPGOS_OK && cudaGetDeviceProperties(&prop, d.Device) != cudaSuccess) rc = fail(MPGOS_ERR_CUDA, "cudaGetDeviceProperties failed");
	if (rc == MPGOS_OK)
	{
		s->sm_count = prop.multiProcessorCount;
		auto ok = [&](cudaError_t e, const char* what) { if (rc == MPGOS_OK && e != cudaSuccess) rc = fail(MPGOS_ERR_CUDA, "%s -> %s", what, cudaGetErrorString(e)); };
		ok(cudaStreamCreate(&s->stream), "cudaStreamCreate");
		ok(cudaEventCreateWithFlags(&s->event, cudaEventDisableTiming), "cudaEventCreate");
		ok(cudaEventCreate(&s->t0), "cudaEventCreate");
		ok(cudaEventCreate(&s->t1), "cudaEventCreate");
		ok(cudaMalloc(&s->d_attempted, 8 * NT), "cudaMalloc");
		ok(cudaMalloc(&s->d_accepted, 8 * NT), "cudaMalloc");
		ok(cudaMalloc(&s->d_totals, 16), "cudaMalloc");
		ok(cudaMalloc(&s->d_queue, 4), "cudaMalloc");
		ok(cudaMalloc(&s->d_last_steps, 4 * NT), "cudaMalloc");
		ok(cudaMalloc(&s->d_order, 4 * (NT / 32 + 1)), "cudaMalloc");
		ok(cudaMalloc(&s->d_order_scratch, 4 * (NT / 32 + 1 + 768)), "cudaMalloc");
		if (rc == MPGOS_OK)
		{
			ok(cudaMemset(s->d_last_steps, 0, 4 * NT), "cudaMemset");
			ok(cudaMemset(s->d_attempted, 0, 8 * NT), "cudaMemset");
			ok(cudaMemset(s->d_accepted, 0, 8 * NT), "cudaMemset");
			ok(cudaMemset(s->d_queue, 0, 4), "cudaMemset");
		}
	}
	if (rc != MPGOS_OK) { char keep[512]; std::memcpy(keep, g_error, sizeof(keep)); mpgos_perthread_destroy(s); std::memcpy(g_error, keep, sizeof(keep)); return rc; }
	*out = s;
	return MPGOS_OK;
}

int mpgos_perthread_create(const mpgos_perthread_desc* desc, mpgos_perthread_solver** out) { return create_impl(desc, true, out); }

int mpgos_perthread_create_sharded(const mpgos_perthread_desc* desc, const int* devices, int device_count, int tile, mpgos_perthread_solver** out)
{
	if (!desc || !out || !devices || device_count < 1) return fail(MPGOS_ERR_ARGUMENT, "create_sharded: null argument or no device");
	*out = nullptr;
	if (tile <= 0) tile = 4096;
	if (tile % 32 != 0) return fail(MPGOS_ERR_ARGUMENT, "create_sharded: the tile (%d) must be a multiple of 32", tile);
	const mpgos_perthread_desc& d = *desc;
	if (d.NT <= 0 || d.SD <= 0 || d.NCP < 0 || d.NSP < 0 || d.NISP < 0 || d.NE < 0 || d.NA < 0 || d.NIA < 0 || d.NDO < 0)
		return fail(MPGOS_ERR_ARGUMENT, "invalid sizes in descriptor");
	if (d.PrecisionBytes != 8 && d.PrecisionBytes != 4) return fail(MPGOS_ERR_ARGUMENT, "precision must be 8 or 4 bytes");
	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	for (int i = 0; i < device_count; i++)
		if (devices[i] < 0 || devices[i] >= ndev) return fail(MPGOS_ERR_ARGUMENT, "device %d not present (%d devices)", devices[i], ndev);

	mpgos_perthread_solver* p = new (std::nothrow) mpgos_perthread_solver();
	if (!p) return fail(MPGOS_ERR_MEMORY, "host allocation failed");
	p->desc = d;  p->desc.Device = devices[0];  p->desc.ShardIndex = 0;  p->desc.ShardCount = device_count;  p->desc.ShardTile = tile;
	p->tile = tile;
	p->ActiveNumberOfThreads = d.NT;
	p->rtol.assign(d.SD, 1e-8);  p->atol.assign(d.SD, 1e-8);
	p->evtol.assign(d.NE > 0 ? d.NE : 1, 1e-6);  p->evdir.assign(d.NE > 0 ? d.NE : 1, 0);
	p->shared_values.assign(d.NSP > 0 ? d.NSP : 1, 0.0);  p->integer_shared_values.assign(d.NISP > 0 ? d.NISP : 1, 0);
	// the host side: one pinned mirror per variable, in the reference's layout over the whole sweep
	const size_t P = (size_t)d.PrecisionBytes, NT = (size_t)d.NT;
	int rc = MPGOS_OK;
	auto A = [&](int v, size_t elements, size_t elem) { if (rc == MPGOS_OK) rc = alloc_mirror(p->var[v], elements, elem, true, false); };
	A(MPGOS_VAR_TimeDomain, 2 * NT, P);  A(MPGOS_VAR_ActualState, d.SD * NT, P);  A(MPGOS_VAR_ActualTime, NT, P);
	A(MPGOS_VAR_ControlParameters, d.NCP * NT, P);  A(MPGOS_VAR_SharedParameters, d.NSP, P);  A(MPGOS_VAR_IntegerSharedParameters, d.NISP, 4);
	A(MPGOS_VAR_Accessories, d.NA * NT, P);  A(MPGOS_VAR_IntegerAccessories, d.NIA * NT, 4);  A(MPGOS_VAR_DenseIndex, NT, 4);
	A(MPGOS_VAR_DenseTime, (size_t)d.NDO * NT, P);  A(MPGOS_VAR_DenseState, (size_t)d.NDO * d.SD * NT, P);
	for (int i = 0; i < device_count && rc == MPGOS_OK; i++)
	{
		mpgos_perthread_desc c = d;
		c.NT = shard_instances(d.NT, tile, device_count, i);
		c.Device = devices[i];
		c.GlobalInstances = d.GlobalInstances > 0 ? d.GlobalInstances : d.NT;
		c.ShardIndex = i;  c.ShardCount = device_count;  c.ShardTile = tile;
		if (c.NT == 0) { rc = fail(MPGOS_ERR_ARGUMENT, "create_sharded: %d instances in tiles of %d leave device %d without work; use fewer devices or a smaller tile", d.NT, tile, devices[i]); break; }
		mpgos_perthread_solver* child = nullptr;
		rc = create_impl(&c, false, &child);
		if (rc == MPGOS_OK) p->shards.push_back(child);
	}
	if (rc != MPGOS_OK) { char keep[512]; std::memcpy(keep, g_error, sizeof(keep)); mpgos_perthread_destroy(p); std::memcpy(g_error, keep, sizeof(keep)); return rc; }
	p->sm_count = p->shards[0]->sm_count;
	*out = p;
	return MPGOS_OK;
}

int mpgos_perthread_shard_count(mpgos_perthread_solver* s) { return !s ? 0 : (s->shards.empty() ? 1 : (int)s->shards.size()); }

int mpgos_perthread_shard_info(mpgos_perthread_solver* s, int shard, int* device, int* instances)
{
	if (!s || shard < 0 || shard >= mpgos_perthread_shard_count(s)) return fail(MPGOS_ERR_ARGUMENT, "shard_info: bad argument");
	mpgos_perthread_solver* c = s->shards.empty() ? s : s->shards[shard];
	if (device) *device = c->desc.Device;
	if (instances) *instances = c->desc.NT;
	return MPGOS_OK;
}

int mpgos_perthread_shard_global_id(mpgos_perthread_solver* s, int shard, int local)
{
	if (!s || shard < 0 || shard >= mpgos_perthread_shard_count(s) || local < 0) return -1;
	if (s->shards.empty()) return local < s->desc.NT ? local + s->desc.InstanceOffset : -1;
	if (local >= s->shards[shard]->desc.NT) return -1;
	return (int)shard_global_id(s->tile, (int)s->shards.size(), shard, local) + s->desc.InstanceOffset;
}

// the map itself, for callers and tests without a device: instances of part `index` of `count`, and a local instance's global id
int mpgos_shard_instances(int NT, int tile, int count, int index) { return (NT < 0 || tile <= 0 || count <= 0 || index < 0 || index >= count) ? -1 : shard_instances(NT, tile, count, index); }
int mpgos_shard_global_id(int tile, int count, int index, int local) { return (tile <= 0 || count <= 0 || index < 0 || index >= count || local < 0) ? -1 : (int)shard_global_id(tile, count, index, local); }

int mpgos_perthread_destroy(mpgos_perthread_solver* s)
{
	if (!s) return MPGOS_OK;
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) mpgos_perthread_destroy(c);
		for (Mirror& m : s->var) if (m.host) cudaFreeHost(m.host);
		delete s;
		return MPGOS_OK;
	}
	cudaSetDevice(s->desc.Device);
	if (s->stream) cudaStreamSynchronize(s->stream);
	for (Mirror& m : s->var) { if (m.host) cudaFreeHost(m.host); if (m.dev) cudaFree(m.dev); }
	for (void* p : s->snapshot) if (p) cudaFree(p);
	if (s->d_attempted) cudaFree(s->d_attempted);
	if (s->d_accepted) cudaFree(s->d_accepted);
	if (s->d_totals) cudaFree(s->d_totals);
	if (s->d_queue) cudaFree(s->d_queue);
	if (s->d_last_steps) cudaFree(s->d_last_steps);
	if (s->d_order) cudaFree(s->d_order);
	if (s->d_order_scratch) cudaFree(s->d_order_scratch);
	if (s->t0) cudaEventDestroy(s->t0);
	if (s->t1) cudaEventDestroy(s->t1);
	if (s->event) cudaEventDestroy(s->event);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
	return MPGOS_OK;
}

static bool is_buffer_variable(int v)
{
	return (v >= MPGOS_VAR_TimeDomain && v <= MPGOS_VAR_IntegerAccessories) || (v >= MPGOS_VAR_DenseIndex && v <= MPGOS_VAR_DenseState);
}

void* mpgos_perthread_host_ptr(mpgos_perthread_solver* s, int variable, size_t* elements)
{
	if (!s || !is_buffer_variable(variable)) { fail(MPGOS_ERR_ARGUMENT, "host_ptr: bad variable %d", variable); return nullptr; }
	if (elements) *elements = s->var[variable].elements;
	return s->var[variable].host;
}

void* mpgos_perthread_device_ptr(mpgos_perthread_solver* s, int variable, size_t* elements)
{
	if (!s || !is_buffer_variable(variable)) { fail(MPGOS_ERR_ARGUMENT, "device_ptr: bad variable %d", variable); return nullptr; }
	if (!s->shards.empty()) return mpgos_perthread_device_ptr(s->shards[0], variable, elements);
	if (elements) *elements = s->var[variable].elements;
	return s->var[variable].dev;
}

int mpgos_perthread_set_option(mpgos_perthread_solver* s, int option, double value)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	for (size_t i = 0; i < s->shards.size(); i++)
	{
		// every part takes the option; "the first n instances are active" becomes each part's share of them
		const double v = option == MPGOS_OPT_ActiveNumberOfThreads ? (double)shard_instances_below(s->desc.NT, s->tile, (int)s->shards.size(), (int)i, (int)value) : value;
		const int rc = mpgos_perthread_set_option(s->shards[i], option, v);
		if (rc != MPGOS_OK) return rc;
	}
	switch (option)
	{
		case MPGOS_OPT_ThreadsPerBlock:            s->ThreadsPerBlock = (int)value; break;
		case MPGOS_OPT_InitialTimeStep:            s->InitialTimeStep = value; break;
		case MPGOS_OPT_ActiveNumberOfThreads:      s->ActiveNumberOfThreads = (int)value; break;
		case MPGOS_OPT_MaximumTimeStep:            s->MaximumTimeStep = value; break;
		case MPGOS_OPT_MinimumTimeStep:            s->MinimumTimeStep = value; break;
		case MPGOS_OPT_TimeStepGrowLimit:          s->TimeStepGrowLimit = value; break;
		case MPGOS_OPT_TimeStepShrinkLimit:        s->TimeStepShrinkLimit = value; break;
		case MPGOS_OPT_DenseOutputMinimumTimeStep: s->DenseOutputMinimumTimeStep = value; break;
		case MPGOS_OPT_DenseOutputSaveFrequency:   s->DenseOutputSaveFrequency = (int)value; break;
		case MPGOS_OPT_PreferSharedMemory:         s->PreferSharedMemory = (int)value; break;
		case MPGOS_OPT_LaneRefill:                 s->LaneRefill = (int)value < 0 ? 0 : ((int)value > 32 ? 32 : (int)value); break;
		case MPGOS_OPT_DeviceMessages:             s->DeviceMessages = (int)value != 0; break;
		case MPGOS_OPT_DispatchOrder:              s->DispatchOrder = (int)value != 0; break;
		default: return fail(MPGOS_ERR_ARGUMENT, "SolverOption: option %d needs 2 input arguments or is not applicable", option);
	}
	return MPGOS_OK;
}

int mpgos_perthread_set_option_indexed(mpgos_perthread_solver* s, int option, int index, double value)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	for (mpgos_perthread_solver* c : s->shards)
	{
		const int rc = mpgos_perthread_set_option_indexed(c, option, index, value);
		if (rc != MPGOS_OK) return rc;
	}
	const int SD = s->desc.SD, NE = s->desc.NE;
	switch (option)
	{
		case MPGOS_OPT_RelativeTolerance:
			if (index < 0 || index >= SD) return fail(MPGOS_ERR_BOUNDS, "RelativeTolerance index %d outside 0-%d", index, SD - 1);
			s->rtol[index] = value; break;
		case MPGOS_OPT_AbsoluteTolerance:
			if (index < 0 || index >= SD) return fail(MPGOS_ERR_BOUNDS, "AbsoluteTolerance index %d outside 0-%d", index, SD - 1);
			s->atol[index] = value; break;
		case MPGOS_OPT_EventTolerance:
			if (index < 0 || index >= NE) return fail(MPGOS_ERR_BOUNDS, "EventTolerance index %d outside 0-%d", index, NE - 1);
			s->evtol[index] = value; break;
		case MPGOS_OPT_EventDirection:
			if (index < 0 || index >= NE) return fail(MPGOS_ERR_BOUNDS, "EventDirection index %d outside 0-%d", index, NE - 1);
			s->evdir[index] = (int)value; break;
		default: return fail(MPGOS_ERR_ARGUMENT, "SolverOption: option %d needs 1 input argument or is not applicable", option);
	}
	return MPGOS_OK;
}

int mpgos_perthread_get_option(mpgos_perthread_solver* s, int option, int index, double* value)
{
	if (!s || !value) return fail(MPGOS_ERR_ARGUMENT, "null argument");
	const int SD = s->desc.SD, NE = s->desc.NE;
	switch (option)
	{
		case MPGOS_OPT_ThreadsPerBlock:            *value = s->ThreadsPerBlock; break;
		case MPGOS_OPT_InitialTimeStep:            *value = s->InitialTimeStep; break;
		case MPGOS_OPT_ActiveNumberOfThreads:      *value = s->ActiveNumberOfThreads; break;
		case MPGOS_OPT_MaximumTimeStep:            *value = s->MaximumTimeStep; break;
		case MPGOS_OPT_MinimumTimeStep:            *value = s->MinimumTimeStep; break;
		case MPGOS_OPT_TimeStepGrowLimit:          *value = s->TimeStepGrowLimit; break;
		case MPGOS_OPT_TimeStepShrinkLimit:        *value = s->TimeStepShrinkLimit; break;
		case MPGOS_OPT_DenseOutputMinimumTimeStep: *value = s->DenseOutputMinimumTimeStep; break;
		case MPGOS_OPT_DenseOutputSaveFrequency:   *value = s->DenseOutputSaveFrequency; break;
		case MPGOS_OPT_PreferSharedMemory:         *value = s->PreferSharedMemory; break;
		case MPGOS_OPT_LaneRefill:                 *value = s->LaneRefill; break;
		case MPGOS_OPT_DeviceMessages:             *value = s->DeviceMessages; break;
		case MPGOS_OPT_DispatchOrder:              *value = s->DispatchOrder; break;
		case MPGOS_OPT_RelativeTolerance: if (index < 0 || index >= SD) return fail(MPGOS_ERR_BOUNDS, "index"); *value = s->rtol[index]; break;
		case MPGOS_OPT_AbsoluteTolerance: if (index < 0 || index >= SD) return fail(MPGOS_ERR_BOUNDS, "index"); *value = s->atol[index]; break;
		case MPGOS_OPT_EventTolerance:    if (index < 0 || index >= NE) return fail(MPGOS_ERR_BOUNDS, "index"); *value = s->evtol[index]; break;
		case MPGOS_OPT_EventDirection:    if (index < 0 || index >= NE) return fail(MPGOS_ERR_BOUNDS, "index"); *value = s->evdir[index]; break;
		default: return fail(MPGOS_ERR_ARGUMENT, "unknown option %d", option);
	}
	return MPGOS_OK;
}

// what the device copy of the shared parameters holds once the copy enqueued now has run
static void shadow_shared(mpgos_perthread_solver* s, int variable, const void* src)
{
	if (variable == MPGOS_VAR_SharedParameters)
		for (int i = 0; i < s->desc.NSP; i++)
			s->shared_values[i] = s->desc.PrecisionBytes == 8 ? ((const double*)src)[i] : (double)((const float*)src)[i];
	if (variable == MPGOS_VAR_IntegerSharedParameters)
		for (int i = 0; i < s->desc.NISP; i++) s->integer_shared_values[i] = ((const int*)src)[i];
}

static int copy_variable(mpgos_perthread_solver* s, int variable, bool to_device)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	// groups exactly as the reference: All = every mirror (Interface.cuh:878-890),
	// DenseOutput = index + time + state (Interface.cuh:924-928)
	int list[11];  int n = 0;
	if (variable == MPGOS_VAR_All)
	{
		for (int v = MPGOS_VAR_TimeDomain; v <= MPGOS_VAR_IntegerAccessories; v++) list[n++] = v;
		list[n++] = MPGOS_VAR_DenseIndex;  list[n++] = MPGOS_VAR_DenseTime;  list[n++] = MPGOS_VAR_DenseState;
	}
	else if (variable == MPGOS_VAR_DenseOutput) { list[n++] = MPGOS_VAR_DenseIndex;  list[n++] = MPGOS_VAR_DenseTime;  list[n++] = MPGOS_VAR_DenseState; }
	else if (is_buffer_variable(variable)) list[n++] = variable;
	else return fail(MPGOS_ERR_ARGUMENT, "Synchronise: bad variable %d", variable);
	if (!s->shards.empty())
	{
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_perthread_solver* c = s->shards[k];
			if (to_device && variable != MPGOS_VAR_TimeDomain && variable != MPGOS_VAR_DenseOutput && variable != MPGOS_VAR_DenseIndex) c->history_valid = false;
			for (int i = 0; i < n; i++)
			{
				Mirror& m = s->var[list[i]];
				if (m.bytes == 0) continue;
				if (serials_of(s->desc, list[i]) > 0)
				{
					const int rc = copy_shard(s, (int)k, list[i], (unsigned char*)m.host, to_device);
					if (rc != MPGOS_OK) return rc;
				}
				else       // shared parameters: every device holds them all
				{
					CUDA_TRY(cudaSetDevice(c->desc.Device));
					if (to_device) { CUDA_TRY(cudaMemcpyAsync(c->var[list[i]].dev, m.host, m.bytes, cudaMemcpyHostToDevice, c->stream));  shadow_shared(c, list[i], m.host); }
					else if (k == 0) CUDA_TRY(cudaMemcpyAsync(m.host, c->var[list[i]].dev, m.bytes, cudaMemcpyDeviceToHost, c->stream));
				}
			}
		}
		return MPGOS_OK;
	}
	// new instances or parameters: the cost history no longer describes them (a new time domain alone leaves the instances what they were)
	if (to_device && variable != MPGOS_VAR_TimeDomain && variable != MPGOS_VAR_DenseOutput && variable != MPGOS_VAR_DenseIndex) s->history_valid = false;
	for (int i = 0; i < n; i++)
	{
		Mirror& m = s->var[list[i]];
		if (m.bytes == 0) continue;
		if (to_device) { CUDA_TRY(cudaMemcpyAsync(m.dev, m.host, m.bytes, cudaMemcpyHostToDevice, s->stream));  shadow_shared(s, list[i], m.host); }
		else           CUDA_TRY(cudaMemcpyAsync(m.host, m.dev, m.bytes, cudaMemcpyDeviceToHost, s->stream));
	}
	return MPGOS_OK;
}

int mpgos_perthread_sync_h2d(mpgos_perthread_solver* s, int variable) { return copy_variable(s, variable, true); }
int mpgos_perthread_sync_d2h(mpgos_perthread_solver* s, int variable) { return copy_variable(s, variable, false); }

int mpgos_perthread_upload(mpgos_perthread_solver* s, int variable, const void* host_src, size_t bytes)
{
	if (!s || !is_buffer_variable(variable) || (!host_src && bytes)) return fail(MPGOS_ERR_ARGUMENT, "upload: bad argument");
	Mirror& m = s->var[variable];
	if (bytes != m.bytes) return fail(MPGOS_ERR_BOUNDS, "upload: variable %d holds %zu bytes, caller passed %zu", variable, m.bytes, bytes);
	if (bytes == 0) return MPGOS_OK;
	if (!s->shards.empty())
	{
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_perthread_solver* c = s->shards[k];
			c->history_valid = false;
			if (serials_of(s->desc, variable) > 0) { const int rc = copy_shard(s, (int)k, variable, (unsigned char*)const_cast<void*>(host_src), true);  if (rc != MPGOS_OK) return rc; }
			else
			{
				CUDA_TRY(cudaSetDevice(c->desc.Device));
				CUDA_TRY(cudaMemcpyAsync(c->var[variable].dev, host_src, bytes, cudaMemcpyHostToDevice, c->stream));
				shadow_shared(c, variable, host_src);
			}
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaMemcpyAsync(m.dev, host_src, bytes, cudaMemcpyHostToDevice, s->stream));
	shadow_shared(s, variable, host_src);
	s->history_valid = false;
	return MPGOS_OK;
}

int mpgos_perthread_download(mpgos_perthread_solver* s, int variable, void* host_dst, size_t bytes)
{
	if (!s || !is_buffer_variable(variable) || (!host_dst && bytes)) return fail(MPGOS_ERR_ARGUMENT, "download: bad argument");
	Mirror& m = s->var[variable];
	if (bytes != m.bytes) return fail(MPGOS_ERR_BOUNDS, "download: variable %d holds %zu bytes, caller passed %zu", variable, m.bytes, bytes);
	if (bytes == 0) return MPGOS_OK;
	if (!s->shards.empty())
	{
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_perthread_solver* c = s->shards[k];
			if (serials_of(s->desc, variable) > 0) { const int rc = copy_shard(s, (int)k, variable, (unsigned char*)host_dst, false);  if (rc != MPGOS_OK) return rc; }
			else if (k == 0)
			{
				CUDA_TRY(cudaSetDevice(c->desc.Device));
				CUDA_TRY(cudaMemcpyAsync(host_dst, c->var[variable].dev, bytes, cudaMemcpyDeviceToHost, c->stream));
			}
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaMemcpyAsync(host_dst, m.dev, bytes, cudaMemcpyDeviceToHost, s->stream));
	return MPGOS_OK;
}

static void fill_launch(mpgos_perthread_solver* s, mpgos_perthread_launch& L)
{
	L.AbiVersion = MPGOS_B200_ABI_VERSION;
	L.desc = s->desc;
	L.d_TimeDomain = s->var[MPGOS_VAR_TimeDomain].dev;  L.d_ActualState = s->var[MPGOS_VAR_ActualState].dev;
	L.d_ActualTime = s->var[MPGOS_VAR_ActualTime].dev;  L.d_ControlParameters = s->var[MPGOS_VAR_ControlParameters].dev;
	L.d_SharedParameters = s->var[MPGOS_VAR_SharedParameters].dev;
	L.d_IntegerSharedParameters = (int*)s->var[MPGOS_VAR_IntegerSharedParameters].dev;
	L.SharedParameterValues = s->shared_values.data();  L.IntegerSharedParameterValues = s->integer_shared_values.data();
	L.d_Accessories = s->var[MPGOS_VAR_Accessories].dev;  L.d_IntegerAccessories = (int*)s->var[MPGOS_VAR_IntegerAccessories].dev;
	L.d_DenseOutputIndex = (int*)s->var[MPGOS_VAR_DenseIndex].dev;
	L.d_DenseOutputTimeInstances = s->var[MPGOS_VAR_DenseTime].dev;  L.d_DenseOutputStates = s->var[MPGOS_VAR_DenseState].dev;
	L.d_StepsAttempted = s->d_attempted;  L.d_StepsAccepted = s->d_accepted;  L.d_WorkQueue = s->d_queue;
	L.d_LastPhaseSteps = s->d_last_steps;  L.d_DispatchOrder = s->d_order;  L.d_DispatchScratch = s->d_order_scratch;
	L.DispatchHistory = (s->DispatchOrder && s->history_valid) ? 1 : 0;
	L.RelativeTolerance = s->rtol.data();  L.AbsoluteTolerance = s->atol.data();
	L.EventTolerance = s->evtol.data();  L.EventDirection = s->evdir.data();
	L.InitialTimeStep = s->InitialTimeStep;  L.MaximumTimeStep = s->MaximumTimeStep;  L.MinimumTimeStep = s->MinimumTimeStep;
	L.TimeStepGrowLimit = s->TimeStepGrowLimit;  L.TimeStepShrinkLimit = s->TimeStepShrinkLimit;
	L.DenseOutputMinimumTimeStep = s->DenseOutputMinimumTimeStep;  L.DenseOutputSaveFrequency = s->DenseOutputSaveFrequency;
	L.ThreadsPerBlock = s->ThreadsPerBlock;  L.ActiveNumberOfThreads = s->ActiveNumberOfThreads;
	L.PreferSharedMemory = s->PreferSharedMemory;  L.LaneRefill = s->LaneRefill;  L.DeviceMessages = s->DeviceMessages;
	L.MultiProcessorCount = s->sm_count;
}

int mpgos_perthread_solve(mpgos_perthread_solver* s, mpgos_perthread_launcher launcher)
{
	if (!s || !launcher) return fail(MPGOS_ERR_ARGUMENT, "Solve: null solver or launcher");
	for (mpgos_perthread_solver* c : s->shards)      // one launch per device, back to back from this thread; each on its own stream
	{
		const int rc = mpgos_perthread_solve(c, launcher);
		if (rc != MPGOS_OK) return rc;
	}
	if (!s->shards.empty()) return MPGOS_OK;
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	mpgos_perthread_launch L;
	fill_launch(s, L);
	const int e = launcher(&L, (void*)s->stream);
	if (e != 0) return fail(MPGOS_ERR_CUDA, "Solve: kernel launch failed -> %s", cudaGetErrorString((cudaError_t)e));
	s->history_valid = true;
	return MPGOS_OK;
}

int mpgos_perthread_solve_timed(mpgos_perthread_solver* s, mpgos_perthread_launcher launcher, int phases, float* elapsed_ms)
{
	if (!s || !launcher || !elapsed_ms || phases < 0) return fail(MPGOS_ERR_ARGUMENT, "solve_timed: bad argument");
	if (!s->shards.empty())
	{
		// all devices run concurrently; the time of the slowest one
		for (mpgos_perthread_solver* c : s->shards)
		{
			CUDA_TRY(cudaSetDevice(c->desc.Device));
			CUDA_TRY(cudaEventRecord(c->t0, c->stream));
			for (int p = 0; p < phases; p++) { const int rc = mpgos_perthread_solve(c, launcher);  if (rc != MPGOS_OK) return rc; }
			CUDA_TRY(cudaEventRecord(c->t1, c->stream));
		}
		*elapsed_ms = 0;
		for (mpgos_perthread_solver* c : s->shards)
		{
			float ms = 0;
			CUDA_TRY(cudaSetDevice(c->desc.Device));
			CUDA_TRY(cudaEventSynchronize(c->t1));
			CUDA_TRY(cudaEventElapsedTime(&ms, c->t0, c->t1));
			if (ms > *elapsed_ms) *elapsed_ms = ms;
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	mpgos_perthread_launch L;
	fill_launch(s, L);
	CUDA_TRY(cudaEventRecord(s->t0, s->stream));
	for (int p = 0; p < phases; p++)
	{
		L.DispatchHistory = (s->DispatchOrder && s->history_valid) ? 1 : 0;
		const int e = launcher(&L, (void*)s->stream);
		if (e != 0) return fail(MPGOS_ERR_CUDA, "Solve: kernel launch failed -> %s", cudaGetErrorString((cudaError_t)e));
		s->history_valid = true;
	}
	CUDA_TRY(cudaEventRecord(s->t1, s->stream));
	CUDA_TRY(cudaEventSynchronize(s->t1));
	CUDA_TRY(cudaEventElapsedTime(elapsed_ms, s->t0, s->t1));
	return MPGOS_OK;
}

int mpgos_perthread_insert_sync_point(mpgos_perthread_solver* s)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) { const int rc = mpgos_perthread_insert_sync_point(c);  if (rc != MPGOS_OK) return rc; }
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaEventRecord(s->event, s->stream));
	return MPGOS_OK;
}

int mpgos_perthread_sync_solver(mpgos_perthread_solver* s)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) { const int rc = mpgos_perthread_sync_solver(c);  if (rc != MPGOS_OK) return rc; }
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaEventSynchronize(s->event));
	return MPGOS_OK;
}

int mpgos_perthread_sync_device(mpgos_perthread_solver* s)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) { const int rc = mpgos_perthread_sync_device(c);  if (rc != MPGOS_OK) return rc; }
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaDeviceSynchronize());
	return MPGOS_OK;
}

int mpgos_perthread_read_step_counters(mpgos_perthread_solver* s, unsigned long long* attempted, unsigned long long* accepted)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		// each part's counters, scattered to their places in the sweep
		std::vector<unsigned long long> a, b;
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_perthread_solver* c = s->shards[k];
			a.resize(c->desc.NT);  b.resize(c->desc.NT);
			const int rc = mpgos_perthread_read_step_counters(c, attempted ? a.data() : nullptr, accepted ? b.data() : nullptr);
			if (rc != MPGOS_OK) return rc;
			for (int i = 0; i < c->desc.NT; i++)
			{
				const long long g = shard_global_id(s->tile, (int)s->shards.size(), (int)k, i);
				if (attempted) attempted[g] = a[i];
				if (accepted)  accepted[g] = b[i];
			}
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	const size_t bytes = 8 * (size_t)s->desc.NT;
	if (attempted) CUDA_TRY(cudaMemcpyAsync(attempted, s->d_attempted, bytes, cudaMemcpyDeviceToHost, s->stream));
	if (accepted)  CUDA_TRY(cudaMemcpyAsync(accepted, s->d_accepted, bytes, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	return MPGOS_OK;
}

int mpgos_perthread_reset_step_counters(mpgos_perthread_solver* s)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) { const int rc = mpgos_perthread_reset_step_counters(c);  if (rc != MPGOS_OK) return rc; }
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	const size_t bytes = 8 * (size_t)s->desc.NT;
	CUDA_TRY(cudaMemsetAsync(s->d_attempted, 0, bytes, s->stream));
	CUDA_TRY(cudaMemsetAsync(s->d_accepted, 0, bytes, s->stream));
	return MPGOS_OK;
}

int mpgos_perthread_total_steps(mpgos_perthread_solver* s, unsigned long long* attempted, unsigned long long* accepted)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		unsigned long long ta = 0, tb = 0;
		for (mpgos_perthread_solver* c : s->shards)
		{
			unsigned long long a = 0, b = 0;
			const int rc = mpgos_perthread_total_steps(c, &a, &b);
			if (rc != MPGOS_OK) return rc;
			ta += a;  tb += b;
		}
		if (attempted) *attempted = ta;
		if (accepted)  *accepted = tb;
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaMemsetAsync(s->d_totals, 0, 16, s->stream));
	reduce_step_counters<<<148 * 4, 256, 0, s->stream>>>(s->d_attempted, s->d_accepted, s->desc.NT, s->d_totals);
	CUDA_TRY(cudaGetLastError());
	unsigned long long h[2] = {0, 0};
	CUDA_TRY(cudaMemcpyAsync(h, s->d_totals, 16, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	if (attempted) *attempted = h[0];
	if (accepted)  *accepted = h[1];
	return MPGOS_OK;
}

static const int kStateVariables[] = { MPGOS_VAR_TimeDomain, MPGOS_VAR_ActualState, MPGOS_VAR_ActualTime, MPGOS_VAR_ControlParameters,
                                       MPGOS_VAR_Accessories, MPGOS_VAR_IntegerAccessories, MPGOS_VAR_DenseIndex };

int mpgos_perthread_save_state(mpgos_perthread_solver* s)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) { const int rc = mpgos_perthread_save_state(c);  if (rc != MPGOS_OK) return rc; }
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	for (int v : kStateVariables)
	{
		Mirror& m = s->var[v];
		if (m.bytes == 0) continue;
		if (!s->snapshot[v]) CUDA_TRY(cudaMalloc(&s->snapshot[v], m.bytes));
		CUDA_TRY(cudaMemcpyAsync(s->snapshot[v], m.dev, m.bytes, cudaMemcpyDeviceToDevice, s->stream));
	}
	return MPGOS_OK;
}

int mpgos_perthread_restore_state(mpgos_perthread_solver* s)
{
	if (!s) return fail(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		for (mpgos_perthread_solver* c : s->shards) { const int rc = mpgos_perthread_restore_state(c);  if (rc != MPGOS_OK) return rc; }
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	for (int v : kStateVariables)
	{
		Mirror& m = s->var[v];
		if (m.bytes == 0) continue;
		if (!s->snapshot[v]) return fail(MPGOS_ERR_ARGUMENT, "restore_state: no snapshot saved");
		CUDA_TRY(cudaMemcpyAsync(m.dev, s->snapshot[v], m.bytes, cudaMemcpyDeviceToDevice, s->stream));
	}
	s->history_valid = false;      // another trajectory from here on: the step counts of the last phase say nothing about it
	return MPGOS_OK;
}

void* mpgos_perthread_stream(mpgos_perthread_solver* s) { return !s ? nullptr : (s->shards.empty() ? (void*)s->stream : (void*)s->shards[0]->stream); }
int   mpgos_perthread_device(mpgos_perthread_solver* s) { return s ? s->desc.Device : -1; }

int mpgos_probe_fma_peak(int device, int precision_bytes, double* tflops)
{
	if (!tflops || (precision_bytes != 8 && precision_bytes != 4)) return fail(MPGOS_ERR_ARGUMENT, "probe_fma_peak: bad argument");
	CUDA_TRY(cudaSetDevice(device));
	return precision_bytes == 8 ? probe_fma_peak<double>(tflops) : probe_fma_peak<float>(tflops);
}

// ---- accuracy probe of include/mpgos_b200/fast_math.cuh (test support): out0/out1 = the fast functions, ref0/ref1 = CUDA's ----
__global__ void fast_math_probe_kernel(int function, int n, const double* __restrict__ x, const double* __restrict__ y,
                                       double* __restrict__ out0, double* __restrict__ out1, double* __restrict__ ref0, double* __restrict__ ref1)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double a = 0, b = 0, c = 0, d = 0;
	if (function == 0)      { mpgos_b200::fm::sincos(x[i], &a, &b);    ::sincos(x[i], &c, &d); }
	else if (function == 1) { mpgos_b200::fm::sincospi(x[i], &a, &b);  ::sincospi(x[i], &c, &d); }
	else if (function == 2) { a = mpgos_b200::fm::pow(x[i], y[i]);     c = ::pow(x[i], y[i]); }
	else if (function == 6) { mpgos_b200::fm::rcp2(x[i], y[i], &a, &b);  c = 1.0 / x[i];  d = 1.0 / y[i]; }
	else if (function == 7) { a = ::sin(x[i]);  b = ::cos(x[i]);  ::sincos(x[i], &c, &d); }      // the CUDA library against itself: sin, cos | sincos
	else
	{
		// the bundled subroutine: 3 = its pow (x^y; trigonometric arguments in range), 4 = its sincospi(x), 5 = its sincos(y)
		double p, s1, c1, s2, c2;
		if (function == 3) mpgos_b200::fm::pow_sincospi_sincos(x[i], y[i], 0.3 * y[i], 1.7 * y[i], &p, &s1, &c1, &s2, &c2);
		else               mpgos_b200::fm::pow_sincospi_sincos(1.0 + 0.5 * ::fabs(::sin(x[i])), 4.2, x[i], y[i], &p, &s1, &c1, &s2, &c2);
		if (function == 3)      { a = p;   c = ::pow(x[i], y[i]); }
		else if (function == 4) { a = s1;  b = c1;  ::sincospi(x[i], &c, &d); }
		else                    { a = s2;  b = c2;  ::sincos(y[i], &c, &d); }
	}
	out0[i] = a;  out1[i] = b;  ref0[i] = c;  ref1[i] = d;
}

int mpgos_fast_math_probe(int device, int function, int n, const double* x, const double* y, double* out0, double* out1, double* ref0, double* ref1)
{
	if (function < 0 || function > 7 || n < 0 || !x || !y || !out0 || !out1 || !ref0 || !ref1) return fail(MPGOS_ERR_ARGUMENT, "fast_math_probe: bad argument");
	if (n == 0) return MPGOS_OK;
	CUDA_TRY(cudaSetDevice(device));
	double* d = nullptr;
	const size_t bytes = sizeof(double) * (size_t)n;
	CUDA_TRY(cudaMalloc(&d, 6 * bytes));
	CUDA_TRY(cudaMemcpy(d, x, bytes, cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(d + n, y, bytes, cudaMemcpyHostToDevice));
	fast_math_probe_kernel<<<(n + 255) / 256, 256>>>(function, n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, d + 5 * (size_t)n);
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaMemcpy(out0, d + 2 * (size_t)n, bytes, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(out1, d + 3 * (size_t)n, bytes, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(ref0, d + 4 * (size_t)n, bytes, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(ref1, d + 5 * (size_t)n, bytes, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) return fail(MPGOS_ERR_CUDA, "fast_math_probe -> %s", cudaGetErrorString(e));
	return MPGOS_OK;
}

int mpgos_probe_dmma_peak(int device, double* tflops)
{
	if (!tflops) return fail(MPGOS_ERR_ARGUMENT, "probe_dmma_peak: bad argument");
	CUDA_TRY(cudaSetDevice(device));
	return probe_dmma_peak(tflops);
}

// ---- built-in systems -------------------------------------------------------------------------

static void collect(std::vector<mpgos_b200::BuiltinEntry>& all)
{
	int n = 0;
	const mpgos_b200::BuiltinEntry* e;
	e = mpgos_b200::builtin_entries_lorenz(&n);        all.insert(all.end(), e, e + n);
	e = mpgos_b200::builtin_entries_duffing(&n);       all.insert(all.end(), e, e + n);
	e = mpgos_b200::builtin_entries_impact(&n);        all.insert(all.end(), e, e + n);
	e = mpgos_b200::builtin_entries_kellermiksis(&n);  all.insert(all.end(), e, e + n);
}

static const std::vector<mpgos_b200::BuiltinEntry>& entries()
{
	// function-local static with an initialiser: constructed once, thread-safe (C++11)
	static const std::vector<mpgos_b200::BuiltinEntry> all = [] { std::vector<mpgos_b200::BuiltinEntry> v; collect(v); return v; }();
	return all;
}

static const char* const kBuiltinNames[] = { "lorenz", "poincare", "duffing", "impact", "kellermiksis", "kellermiksis_dense", "tidprobe", "kellermiksis_fm", "kellermiksis_dense_fm" };

int mpgos_perthread_builtin_count(void) { return (int)(sizeof(kBuiltinNames) / sizeof(kBuiltinNames[0])); }
const char* mpgos_perthread_builtin_name(int i) { return (i >= 0 && i < mpgos_perthread_builtin_count()) ? kBuiltinNames[i] : nullptr; }

int mpgos_perthread_builtin_desc(const char* system, mpgos_perthread_desc* desc)
{
	if (!system || !desc) return fail(MPGOS_ERR_ARGUMENT, "null argument");
	for (const mpgos_b200::BuiltinEntry& e : entries())
		if (!std::strcmp(e.name, system))
		{
			desc->SD = e.SD; desc->NCP = e.NCP; desc->NSP = e.NSP; desc->NISP = e.NISP;
			desc->NE = e.NE; desc->NA = e.NA; desc->NIA = e.NIA;
			return MPGOS_OK;
		}
	return fail(MPGOS_ERR_UNKNOWN_SYSTEM, "no built-in system named '%s'", system);
}

mpgos_perthread_launcher mpgos_perthread_builtin_launcher(const char* system, int algorithm, int precision_bytes, int dense_output)
{
	if (!system) return nullptr;
	for (const mpgos_b200::BuiltinEntry& e : entries())
		if (!std::strcmp(e.name, system) && e.Algorithm == algorithm && e.PrecisionBytes == precision_bytes && e.Dense == (dense_output != 0))
			return e.launcher;
	fail(MPGOS_ERR_UNKNOWN_SYSTEM, "no built-in launcher for system '%s' algorithm %d precision %d dense %d", system, algorithm, precision_bytes, dense_output);
	return nullptr;
}

} // extern "C"
