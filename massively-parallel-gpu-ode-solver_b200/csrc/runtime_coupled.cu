// runtime_coupled.cu -- C-ABI of the block-per-system (coupled) module, include/mpgos_b200_coupled.h.
// Host half of the reference's coupled ProblemSolver (SourceCodes/CoupledSystems_PerBlock_Interface.cuh:353-1992):
// stream + event, pinned mirrors, device buffers, coupling-matrix packing, options, asynchronous copies, launch.
#include "../../include/mpgos_b200_coupled.h"
#include "builtin_registry.h"

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

extern "C" const char* mpgos_last_error(void);
namespace mpgos_b200 { int set_error(int code, const char* fmt, ...); }   // runtime.cu

#define CFAIL(...) mpgos_b200::set_error(__VA_ARGS__)
#define CUDA_TRY(call)                                                                                   \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess)                                                                           \
			return CFAIL(MPGOS_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

namespace {
struct CMirror { void* host = nullptr; void* dev = nullptr; void* snap = nullptr; size_t elements = 0, bytes = 0; };
}

struct mpgos_coupled_solver
{
	mpgos_coupled_desc desc;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t event = nullptr, t0 = nullptr, t1 = nullptr;
	CMirror var[MPGOS_CVAR_COUNT];
	std::vector<double> original_matrix;     // h_CouplingMatrixOriginal (host only, CP/Interface.cuh:461)
	unsigned long long *d_attempted = nullptr, *d_accepted = nullptr;
	unsigned int* d_queue = nullptr;
	unsigned int* d_last_steps = nullptr;      // [NS] steps of the previous phase (cost history of the queue order)
	int*          d_order = nullptr;           // [NS]
	unsigned int* d_order_scratch = nullptr;   // [NS + 768]
	void*         d_slice = nullptr;           // [NS x 40 bytes] pilot-slice state
	void*         d_slice_params = nullptr;    // parameters as of the pause (allocated on the first phase that uses a pilot slice)
	bool          history_valid = false;
	int           DispatchOrder = 1;
	double InitialTimeStep = 1e-2, MaximumTimeStep = 1.0e6, MinimumTimeStep = 1.0e-12, TimeStepGrowLimit = 5.0, TimeStepShrinkLimit = 0.1;
	double DenseOutputMinimumTimeStep = 0.0;
	int DenseOutputSaveFrequency = 1, ActiveSystems = 0, DeviceMessages = 0, SharedGlobalVariables = 1, SharedCouplingMatrices = 0, TensorCoreCoupling = 1;
	std::vector<double> rtol, atol, evtol;
	std::vector<int> evdir;
	// multi-device solver (mpgos_coupled_create_sharded): the parent owns the host mirrors over all systems and no device memory
	std::vector<mpgos_coupled_solver*> shards;
	int tile = 0;
};

namespace {
// tile-cyclic shard map, as in runtime.cu
int cshard_systems(int NS, int tile, int count, int index)
{
	const int ntiles = (NS + tile - 1) / tile;
	if (index >= ntiles) return 0;
	int n = ((ntiles - index + count - 1) / count) * tile;
	if ((ntiles - 1) % count == index) n -= ntiles * tile - NS;
	return n;
}
long long cshard_global_id(int tile, int count, int index, int local) { return ((long long)(local / tile) * count + index) * tile + local % tile; }
}

static size_t matrix_elements(const mpgos_coupled_desc& d)
{
	if (d.CCI == 0 && d.CBW == 0) return (size_t)d.NC * d.UPS * d.UPS;
	if (d.CCI == 0) return (size_t)d.NC * d.UPS * (2 * d.CBW + 1);
	if (d.CBW > 0) return (size_t)d.NC * (2 * d.CBW + 1);
	return (size_t)d.NC * d.UPS;
}

static bool is_cbuffer(int v) { return v > MPGOS_CVAR_All && v < MPGOS_CVAR_COUNT && v != MPGOS_CVAR_DenseOutput; }
static bool is_int_variable(int v)
{
	return v == MPGOS_CVAR_IntegerGlobalParameters || v == MPGOS_CVAR_IntegerUnitAccessories || v == MPGOS_CVAR_IntegerSystemAccessories ||
	       v == MPGOS_CVAR_CouplingIndex || v == MPGOS_CVAR_DenseIndex;
}

// (elements per system in one serial, serials) of a per-system variable; width 0: replicated on every device
static void scope_of(const mpgos_coupled_desc& d, int v, int* width, int* serials)
{
	*width = 0;  *serials = 0;
	switch (v)
	{
		case MPGOS_CVAR_ActualState:             *width = d.UPS; *serials = d.UD; break;
		case MPGOS_CVAR_UnitParameters:          *width = d.UPS; *serials = d.NUP; break;
		case MPGOS_CVAR_UnitAccessories:         *width = d.UPS; *serials = d.NUA; break;
		case MPGOS_CVAR_IntegerUnitAccessories:  *width = d.UPS; *serials = d.NiUA; break;
		case MPGOS_CVAR_DenseState:              *width = d.UPS; *serials = d.UD * d.NDO; break;
		case MPGOS_CVAR_TimeDomain:              *width = 1; *serials = 2; break;
		case MPGOS_CVAR_ActualTime:              *width = 1; *serials = 1; break;
		case MPGOS_CVAR_SystemParameters:        *width = 1; *serials = d.NSP; break;
		case MPGOS_CVAR_SystemAccessories:       *width = 1; *serials = d.NSA; break;
		case MPGOS_CVAR_IntegerSystemAccessories:*width = 1; *serials = d.NiSA; break;
		case MPGOS_CVAR_CouplingStrength:        *width = 1; *serials = d.NC; break;
		case MPGOS_CVAR_DenseIndex:              *width = 1; *serials = 1; break;
		case MPGOS_CVAR_DenseTime:               *width = 1; *serials = d.NDO; break;
		default: break;
	}
}

// strided copy between a host array over all NS systems and one part's device buffer (see copy_shard in runtime.cu)
static int ccopy_shard(mpgos_coupled_solver* parent, int index, int v, unsigned char* host, bool to_device)
{
	mpgos_coupled_solver* c = parent->shards[index];
	const mpgos_coupled_desc& pd = parent->desc;
	const int count = (int)parent->shards.size(), tile = parent->tile, NS = pd.NS, n = c->desc.NS;
	int width = 0, serials = 0;
	scope_of(pd, v, &width, &serials);
	const size_t es = (is_int_variable(v) ? 4 : (size_t)pd.PrecisionBytes) * (size_t)width;      // bytes per system and serial
	unsigned char* dev = (unsigned char*)c->var[v].dev;
	if (n == 0 || serials == 0 || width == 0 || !dev) return MPGOS_OK;
	CUDA_TRY(cudaSetDevice(c->desc.Device));
	auto copy2d = [&](size_t dev_off, size_t host_off, size_t w, size_t rows) -> cudaError_t
	{
		if (rows == 0 || w == 0) return cudaSuccess;
		return to_device ? cudaMemcpy2DAsync(dev + dev_off, (size_t)tile * es, host + host_off, (size_t)tile * count * es, w, rows, cudaMemcpyHostToDevice, c->stream)
		                 : cudaMemcpy2DAsync(host + host_off, (size_t)tile * count * es, dev + dev_off, (size_t)tile * es, w, rows, cudaMemcpyDeviceToHost, c->stream);
	};
	const size_t first = (size_t)index * tile * es;
	if (NS % (tile * count) == 0) { CUDA_TRY(copy2d(0, first, (size_t)tile * es, (size_t)serials * (n / tile)));  return MPGOS_OK; }
	const int full = n / tile, rest = n % tile;
	for (int k = 0; k < serials; k++)
	{
		CUDA_TRY(copy2d((size_t)k * n * es, (size_t)k * NS * es + first, (size_t)tile * es, (size_t)full));
		if (rest) CUDA_TRY(copy2d(((size_t)k * n + (size_t)full * tile) * es, (size_t)k * NS * es + first + (size_t)full * tile * count * es, (size_t)rest * es, 1));
	}
	return MPGOS_OK;
}

static int ccreate_impl(const mpgos_coupled_desc* desc, bool host_mirrors, bool device_side, mpgos_coupled_solver** out);

extern "C" {

int mpgos_coupled_create(const mpgos_coupled_desc* desc, mpgos_coupled_solver** out) { return ccreate_impl(desc, true, true, out); }

int mpgos_coupled_create_sharded(const mpgos_coupled_desc* desc, const int* devices, int device_count, int tile, mpgos_coupled_solver** out)
{
	if (!desc || !out || !devices || device_count < 1) return CFAIL(MPGOS_ERR_ARGUMENT, "create_sharded: null argument or no device");
	*out = nullptr;
	if (tile <= 0) tile = 64;
	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	for (int i = 0; i < device_count; i++)
		if (devices[i] < 0 || devices[i] >= ndev) return CFAIL(MPGOS_ERR_ARGUMENT, "device %d not present (%d devices)", devices[i], ndev);
	mpgos_coupled_desc pd = *desc;
	pd.Device = devices[0];  pd.ShardIndex = 0;  pd.ShardCount = device_count;  pd.ShardTile = tile;
	mpgos_coupled_solver* p = nullptr;
	int rc = ccreate_impl(&pd, true, false, &p);      // host mirrors over all systems, nothing on a device
	if (rc != MPGOS_OK) return rc;
	p->tile = tile;
	for (int i = 0; i < device_count && rc == MPGOS_OK; i++)
	{
		mpgos_coupled_desc c = *desc;
		c.NS = cshard_systems(desc->NS, tile, device_count, i);
		c.Device = devices[i];  c.ShardIndex = i;  c.ShardCount = device_count;  c.ShardTile = tile;
		if (c.NS == 0) { rc = CFAIL(MPGOS_ERR_ARGUMENT, "create_sharded: %d systems in tiles of %d leave device %d without work; use fewer devices or a smaller tile", desc->NS, tile, devices[i]); break; }
		mpgos_coupled_solver* child = nullptr;
		rc = ccreate_impl(&c, false, true, &child);
		if (rc == MPGOS_OK) p->shards.push_back(child);
	}
	if (rc != MPGOS_OK) { char keep[512]; std::strncpy(keep, mpgos_last_error(), sizeof(keep) - 1); keep[sizeof(keep) - 1] = 0; mpgos_coupled_destroy(p); CFAIL(rc, "%s", keep); return rc; }
	p->sm_count = p->shards[0]->sm_count;
	*out = p;
	return MPGOS_OK;
}

int mpgos_coupled_shard_count(mpgos_coupled_solver* s) { return !s ? 0 : (s->shards.empty() ? 1 : (int)s->shards.size()); }

} // extern "C"

static int ccreate_impl(const mpgos_coupled_desc* desc, bool host_mirrors, bool device_side, mpgos_coupled_solver** out)
{
	if (!desc || !out) return CFAIL(MPGOS_ERR_ARGUMENT, "null argument");
	*out = nullptr;
	const mpgos_coupled_desc& d = *desc;
	if (d.NS <= 0 || d.UPS <= 0 || d.UD <= 0 || d.NC < 0 || d.CBW < 0 || (d.CCI != 0 && d.CCI != 1) || d.NUP < 0 || d.NSP < 0 || d.NGP < 0 || d.NiGP < 0 ||
	    d.NUA < 0 || d.NiUA < 0 || d.NSA < 0 || d.NiSA < 0 || d.NE < 0 || d.NDO < 0)
		return CFAIL(MPGOS_ERR_ARGUMENT, "invalid sizes in coupled descriptor");
	if (d.Algorithm != MPGOS_RK4 && d.Algorithm != MPGOS_RKCK45) return CFAIL(MPGOS_ERR_ARGUMENT, "unknown algorithm %d", d.Algorithm);
	if (d.PrecisionBytes != 8 && d.PrecisionBytes != 4) return CFAIL(MPGOS_ERR_ARGUMENT, "PrecisionBytes must be 8 (double) or 4 (float), got %d", d.PrecisionBytes);
	if (2 * d.CBW + 1 > d.UPS) return CFAIL(MPGOS_ERR_ARGUMENT, "coupling bandwidth %d too wide for %d units", d.CBW, d.UPS);
	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	if (d.Device < 0 || d.Device >= ndev) return CFAIL(MPGOS_ERR_ARGUMENT, "device %d not present (%d devices)", d.Device, ndev);
	CUDA_TRY(cudaSetDevice(d.Device));

	const size_t P = (size_t)d.PrecisionBytes, NS = d.NS, NU = NS * d.UPS;
	size_t elems[MPGOS_CVAR_COUNT] = {0};
	elems[MPGOS_CVAR_TimeDomain] = 2 * NS;  elems[MPGOS_CVAR_ActualTime] = NS;  elems[MPGOS_CVAR_ActualState] = NU * d.UD;
	elems[MPGOS_CVAR_UnitParameters] = NU * d.NUP;  elems[MPGOS_CVAR_SystemParameters] = NS * d.NSP;  elems[MPGOS_CVAR_GlobalParameters] = d.NGP;
	elems[MPGOS_CVAR_IntegerGlobalParameters] = d.NiGP;  elems[MPGOS_CVAR_UnitAccessories] = NU * d.NUA;  elems[MPGOS_CVAR_IntegerUnitAccessories] = NU * d.NiUA;
	elems[MPGOS_CVAR_SystemAccessories] = NS * d.NSA;  elems[MPGOS_CVAR_IntegerSystemAccessories] = NS * d.NiSA;  elems[MPGOS_CVAR_CouplingMatrix] = matrix_elements(d);
	elems[MPGOS_CVAR_CouplingStrength] = NS * d.NC;  elems[MPGOS_CVAR_CouplingIndex] = d.NC;  elems[MPGOS_CVAR_DenseIndex] = NS;
	elems[MPGOS_CVAR_DenseTime] = NS * d.NDO;  elems[MPGOS_CVAR_DenseState] = NU * d.UD * d.NDO;
	size_t need = 16 * NS + 64;
	for (int v = 0; v < MPGOS_CVAR_COUNT; v++) need += elems[v] * (is_int_variable(v) ? 4 : P);
	size_t free_b = 0, total_b = 0;
	CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
	if (device_side && need >= free_b) return CFAIL(MPGOS_ERR_MEMORY, "solver needs %zu bytes of device memory, %zu free", need, free_b);

	mpgos_coupled_solver* s = new (std::nothrow) mpgos_coupled_solver();
	if (!s) return CFAIL(MPGOS_ERR_MEMORY, "host allocation failed");
	s->desc = d;  s->ActiveSystems = d.NS;
	s->rtol.assign(d.UD, 1e-8);  s->atol.assign(d.UD, 1e-8);  s->evtol.assign(d.NE > 0 ? d.NE : 1, 1e-6);  s->evdir.assign(d.NE > 0 ? d.NE : 1, 0);
	s->original_matrix.assign((size_t)d.NC * d.UPS * d.UPS, 0.0);
	cudaError_t e = cudaSuccess;
	for (int v = 0; v < MPGOS_CVAR_COUNT && e == cudaSuccess; v++)
	{
		CMirror& m = s->var[v];
		m.elements = elems[v];  m.bytes = elems[v] * (is_int_variable(v) ? 4 : P);
		if (!m.bytes) continue;
		if (host_mirrors) { e = cudaMallocHost(&m.host, m.bytes);  if (e == cudaSuccess) std::memset(m.host, 0, m.bytes); }
		if (e == cudaSuccess && device_side) e = cudaMalloc(&m.dev, m.bytes);
		if (e == cudaSuccess && device_side) e = cudaMemset(m.dev, 0, m.bytes);
	}
	if (!device_side)
	{
		if (e != cudaSuccess) { mpgos_coupled_destroy(s); return CFAIL(MPGOS_ERR_CUDA, "coupled solver construction failed -> %s", cudaGetErrorString(e)); }
		*out = s;
		return MPGOS_OK;
	}
	cudaDeviceProp prop;
	if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, d.Device);
	if (e == cudaSuccess) { s->sm_count = prop.multiProcessorCount; e = cudaStreamCreate(&s->stream); }
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->event, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreate(&s->t0);
	if (e == cudaSuccess) e = cudaEventCreate(&s->t1);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_attempted, 8 * NS);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_accepted, 8 * NS);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_queue, 4);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_last_steps, 4 * NS);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_order, 4 * NS);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_order_scratch, 4 * (NS + 768));
	if (e == cudaSuccess) e = cudaMemset(s->d_last_steps, 0, 4 * NS);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_slice, 40 * NS);
	if (e == cudaSuccess) e = cudaMemset(s->d_slice, 0, 40 * NS);
	if (e == cudaSuccess) e = cudaMemset(s->d_attempted, 0, 8 * NS);
	if (e == cudaSuccess) e = cudaMemset(s->d_accepted, 0, 8 * NS);
	if (e != cudaSuccess) { mpgos_coupled_destroy(s); return CFAIL(MPGOS_ERR_CUDA, "coupled solver construction failed -> %s", cudaGetErrorString(e)); }
	*out = s;
	return MPGOS_OK;
}

extern "C" {

int mpgos_coupled_destroy(mpgos_coupled_solver* s)
{
	if (!s) return MPGOS_OK;
	for (mpgos_coupled_solver* c : s->shards) mpgos_coupled_destroy(c);
	s->shards.clear();
	cudaSetDevice(s->desc.Device);
	if (s->stream) cudaStreamSynchronize(s->stream);
	for (CMirror& m : s->var) { if (m.host) cudaFreeHost(m.host); if (m.dev) cudaFree(m.dev); if (m.snap) cudaFree(m.snap); }
	if (s->d_attempted) cudaFree(s->d_attempted);
	if (s->d_accepted) cudaFree(s->d_accepted);
	if (s->d_queue) cudaFree(s->d_queue);
	if (s->d_last_steps) cudaFree(s->d_last_steps);
	if (s->d_order) cudaFree(s->d_order);
	if (s->d_order_scratch) cudaFree(s->d_order_scratch);
	if (s->d_slice) cudaFree(s->d_slice);
	if (s->d_slice_params) cudaFree(s->d_slice_params);
	if (s->t0) cudaEventDestroy(s->t0);
	if (s->t1) cudaEventDestroy(s->t1);
	if (s->event) cudaEventDestroy(s->event);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
	return MPGOS_OK;
}

void* mpgos_coupled_host_ptr(mpgos_coupled_solver* s, int variable, size_t* elements)
{
	if (!s || !is_cbuffer(variable)) { CFAIL(MPGOS_ERR_ARGUMENT, "host_ptr: bad variable %d", variable); return nullptr; }
	if (elements) *elements = s->var[variable].elements;
	return s->var[variable].host;
}

// index of (coupling, row, col) in the packed matrix, or -1 when the entry lies outside the stored band
static long packed_index(const mpgos_coupled_desc& d, int k, int row, int col)
{
	const int UPS = d.UPS;
	if (d.CCI == 0 && d.CBW == 0) return (long)row + (long)col * UPS + (long)k * UPS * UPS;
	int mc = col - row + d.CBW;
	if (mc >= UPS) mc -= UPS;
	if (mc < 0) mc += UPS;
	const int lim = d.CBW > 0 ? 2 * d.CBW + 1 : UPS;
	if (mc >= lim) return -1;
	if (d.CCI == 0) return (long)row + (long)mc * UPS + (long)k * UPS * lim;
	return (long)mc + (long)k * lim;
}

int mpgos_coupled_set_matrix_entry(mpgos_coupled_solver* s, int k, int row, int col, double value)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	const mpgos_coupled_desc& d = s->desc;
	if (k < 0 || k >= d.NC) return CFAIL(MPGOS_ERR_BOUNDS, "CouplingMatrix number %d outside 0-%d", k, d.NC - 1);
	if (row < 0 || row >= d.UPS) return CFAIL(MPGOS_ERR_BOUNDS, "Coupling Matrix Row %d outside 0-%d", row, d.UPS - 1);
	if (col < 0 || col >= d.UPS) return CFAIL(MPGOS_ERR_BOUNDS, "Coupling Matrix Col %d outside 0-%d", col, d.UPS - 1);
	const long idx = packed_index(d, k, row, col);
	if (idx < 0) return CFAIL(MPGOS_ERR_BOUNDS, "Coupling Matrix Shifted Diagonal of (%d,%d) outside 0-%d", row, col, (d.CBW > 0 ? 2 * d.CBW + 1 : d.UPS) - 1);
	s->original_matrix[(size_t)row + (size_t)col * d.UPS + (size_t)k * d.UPS * d.UPS] = value;
	if (d.PrecisionBytes == 8) static_cast<double*>(s->var[MPGOS_CVAR_CouplingMatrix].host)[idx] = value;
	else                       static_cast<float*>(s->var[MPGOS_CVAR_CouplingMatrix].host)[idx] = (float)value;
	return MPGOS_OK;
}

int mpgos_coupled_get_matrix_entry(mpgos_coupled_solver* s, int k, int row, int col, double* value)
{
	if (!s || !value) return CFAIL(MPGOS_ERR_ARGUMENT, "null argument");
	const mpgos_coupled_desc& d = s->desc;
	if (k < 0 || k >= d.NC || row < 0 || row >= d.UPS || col < 0 || col >= d.UPS) return CFAIL(MPGOS_ERR_BOUNDS, "coupling matrix index out of range");
	*value = s->original_matrix[(size_t)row + (size_t)col * d.UPS + (size_t)k * d.UPS * d.UPS];
	return MPGOS_OK;
}

int mpgos_coupled_set_option(mpgos_coupled_solver* s, int option, double value)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	for (size_t i = 0; i < s->shards.size(); i++)
	{
		// "the first n systems are active" becomes each part's share of them
		const int lim = (int)value < s->desc.NS ? ((int)value > 0 ? (int)value : 0) : s->desc.NS;
		const double v = option == MPGOS_COPT_ActiveSystems ? (double)cshard_systems(lim, s->tile, (int)s->shards.size(), (int)i) : value;
		const int rc = mpgos_coupled_set_option(s->shards[i], option, v);
		if (rc != MPGOS_OK) return rc;
	}
	switch (option)
	{
		case MPGOS_COPT_InitialTimeStep:            s->InitialTimeStep = value; break;
		case MPGOS_COPT_ActiveSystems:              s->ActiveSystems = (int)value; break;
		case MPGOS_COPT_MaximumTimeStep:            s->MaximumTimeStep = value; break;
		case MPGOS_COPT_MinimumTimeStep:            s->MinimumTimeStep = value; break;
		case MPGOS_COPT_TimeStepGrowLimit:          s->TimeStepGrowLimit = value; break;
		case MPGOS_COPT_TimeStepShrinkLimit:        s->TimeStepShrinkLimit = value; break;
		case MPGOS_COPT_DenseOutputSaveFrequency:   s->DenseOutputSaveFrequency = (int)value; break;
		case MPGOS_COPT_DenseOutputMinimumTimeStep: s->DenseOutputMinimumTimeStep = value; break;
		case MPGOS_COPT_SharedGlobalVariables:      s->SharedGlobalVariables = (int)value; break;     // accepted, ignored: always shared
		case MPGOS_COPT_SharedCouplingMatrices:     s->SharedCouplingMatrices = (int)value; break;    // accepted, ignored: always shared
		case MPGOS_COPT_DeviceMessages:             s->DeviceMessages = (int)value != 0; break;
		case MPGOS_COPT_TensorCoreCoupling:         s->TensorCoreCoupling = (int)value != 0; break;
		case MPGOS_COPT_DispatchOrder:              s->DispatchOrder = (int)value < 0 ? 0 : ((int)value > 2 ? 2 : (int)value); break;
		default: return CFAIL(MPGOS_ERR_ARGUMENT, "SolverOption: option %d needs 2 input arguments or is not applicable", option);
	}
	return MPGOS_OK;
}

int mpgos_coupled_set_option_indexed(mpgos_coupled_solver* s, int option, int index, double value)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_set_option_indexed(c, option, index, value);  if (rc != MPGOS_OK) return rc; }
	const int UD = s->desc.UD, NE = s->desc.NE;
	switch (option)
	{
		case MPGOS_COPT_RelativeTolerance: if (index < 0 || index >= UD) return CFAIL(MPGOS_ERR_BOUNDS, "RelativeTolerance index %d outside 0-%d", index, UD - 1); s->rtol[index] = value; break;
		case MPGOS_COPT_AbsoluteTolerance: if (index < 0 || index >= UD) return CFAIL(MPGOS_ERR_BOUNDS, "AbsoluteTolerance index %d outside 0-%d", index, UD - 1); s->atol[index] = value; break;
		case MPGOS_COPT_EventTolerance:    if (index < 0 || index >= NE) return CFAIL(MPGOS_ERR_BOUNDS, "EventTolerance index %d outside 0-%d", index, NE - 1); s->evtol[index] = value; break;
		case MPGOS_COPT_EventDirection:    if (index < 0 || index >= NE) return CFAIL(MPGOS_ERR_BOUNDS, "EventDirection index %d outside 0-%d", index, NE - 1); s->evdir[index] = (int)value; break;
		default: return CFAIL(MPGOS_ERR_ARGUMENT, "SolverOption: option %d needs 1 input argument or is not applicable", option);
	}
	return MPGOS_OK;
}

static int ccopy(mpgos_coupled_solver* s, int variable, bool to_device)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	int list[MPGOS_CVAR_COUNT];  int n = 0;
	if (variable == MPGOS_CVAR_All) { for (int v = 1; v < MPGOS_CVAR_COUNT; v++) if (is_cbuffer(v)) list[n++] = v; }
	else if (variable == MPGOS_CVAR_DenseOutput) { list[n++] = MPGOS_CVAR_DenseIndex; list[n++] = MPGOS_CVAR_DenseTime; list[n++] = MPGOS_CVAR_DenseState; }
	else if (is_cbuffer(variable)) list[n++] = variable;
	else return CFAIL(MPGOS_ERR_ARGUMENT, "Synchronise: bad variable %d", variable);
	if (!s->shards.empty())
	{
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_coupled_solver* c = s->shards[k];
			if (to_device && variable != MPGOS_CVAR_TimeDomain && variable != MPGOS_CVAR_DenseOutput && variable != MPGOS_CVAR_DenseIndex) c->history_valid = false;
			for (int i = 0; i < n; i++)
			{
				CMirror& m = s->var[list[i]];
				if (!m.bytes) continue;
				int width = 0, serials = 0;
				scope_of(s->desc, list[i], &width, &serials);
				if (width > 0) { const int rc = ccopy_shard(s, (int)k, list[i], (unsigned char*)m.host, to_device);  if (rc != MPGOS_OK) return rc; }
				else       // coupling matrices, global parameters, coupling indices: every device holds them all
				{
					CUDA_TRY(cudaSetDevice(c->desc.Device));
					if (to_device)   CUDA_TRY(cudaMemcpyAsync(c->var[list[i]].dev, m.host, m.bytes, cudaMemcpyHostToDevice, c->stream));
					else if (k == 0) CUDA_TRY(cudaMemcpyAsync(m.host, c->var[list[i]].dev, m.bytes, cudaMemcpyDeviceToHost, c->stream));
				}
			}
		}
		return MPGOS_OK;
	}
	// new systems or parameters: the cost history no longer describes them (a new time domain alone leaves the systems what they were)
	if (to_device && variable != MPGOS_CVAR_TimeDomain && variable != MPGOS_CVAR_DenseOutput && variable != MPGOS_CVAR_DenseIndex) s->history_valid = false;
	for (int i = 0; i < n; i++)
	{
		CMirror& m = s->var[list[i]];
		if (!m.bytes) continue;
		if (to_device) CUDA_TRY(cudaMemcpyAsync(m.dev, m.host, m.bytes, cudaMemcpyHostToDevice, s->stream));
		else           CUDA_TRY(cudaMemcpyAsync(m.host, m.dev, m.bytes, cudaMemcpyDeviceToHost, s->stream));
	}
	return MPGOS_OK;
}
int mpgos_coupled_sync_h2d(mpgos_coupled_solver* s, int variable) { return ccopy(s, variable, true); }
int mpgos_coupled_sync_d2h(mpgos_coupled_solver* s, int variable) { return ccopy(s, variable, false); }

int mpgos_coupled_upload(mpgos_coupled_solver* s, int variable, const void* host_src, size_t bytes)
{
	if (!s || !is_cbuffer(variable) || variable == MPGOS_CVAR_CouplingMatrix || (!host_src && bytes)) return CFAIL(MPGOS_ERR_ARGUMENT, "upload: bad argument");
	CMirror& m = s->var[variable];
	if (bytes != m.bytes) return CFAIL(MPGOS_ERR_BOUNDS, "upload: variable %d holds %zu bytes, caller passed %zu", variable, m.bytes, bytes);
	if (bytes == 0) return MPGOS_OK;
	if (!s->shards.empty())
	{
		int width = 0, serials = 0;
		scope_of(s->desc, variable, &width, &serials);
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_coupled_solver* c = s->shards[k];
			c->history_valid = false;
			if (width > 0) { const int rc = ccopy_shard(s, (int)k, variable, (unsigned char*)const_cast<void*>(host_src), true);  if (rc != MPGOS_OK) return rc; }
			else { CUDA_TRY(cudaSetDevice(c->desc.Device));  CUDA_TRY(cudaMemcpyAsync(c->var[variable].dev, host_src, bytes, cudaMemcpyHostToDevice, c->stream)); }
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaMemcpyAsync(m.dev, host_src, bytes, cudaMemcpyHostToDevice, s->stream));
	s->history_valid = false;
	return MPGOS_OK;
}

int mpgos_coupled_download(mpgos_coupled_solver* s, int variable, void* host_dst, size_t bytes)
{
	if (!s || !is_cbuffer(variable) || variable == MPGOS_CVAR_CouplingMatrix || (!host_dst && bytes)) return CFAIL(MPGOS_ERR_ARGUMENT, "download: bad argument");
	CMirror& m = s->var[variable];
	if (bytes != m.bytes) return CFAIL(MPGOS_ERR_BOUNDS, "download: variable %d holds %zu bytes, caller passed %zu", variable, m.bytes, bytes);
	if (bytes == 0) return MPGOS_OK;
	if (!s->shards.empty())
	{
		int width = 0, serials = 0;
		scope_of(s->desc, variable, &width, &serials);
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_coupled_solver* c = s->shards[k];
			if (width > 0) { const int rc = ccopy_shard(s, (int)k, variable, (unsigned char*)host_dst, false);  if (rc != MPGOS_OK) return rc; }
			else if (k == 0) { CUDA_TRY(cudaSetDevice(c->desc.Device));  CUDA_TRY(cudaMemcpyAsync(host_dst, c->var[variable].dev, bytes, cudaMemcpyDeviceToHost, c->stream)); }
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaMemcpyAsync(host_dst, m.dev, bytes, cudaMemcpyDeviceToHost, s->stream));
	return MPGOS_OK;
}

// the pilot slice pauses systems in mid-phase: room for the parameter arrays as the callbacks left them (allocated on first use)
static int ensure_slice_parameters(mpgos_coupled_solver* s)
{
	if (s->d_slice_params || !s->DispatchOrder) return MPGOS_OK;
	const mpgos_coupled_desc& d = s->desc;
	const size_t elements = ((size_t)d.NUP * d.UPS + d.NSP + d.NC) * (size_t)d.NS;
	CUDA_TRY(cudaMalloc(&s->d_slice_params, (elements ? elements : 1) * (size_t)d.PrecisionBytes));
	return MPGOS_OK;
}

static void cfill(mpgos_coupled_solver* s, mpgos_coupled_launch& L)
{
	L.AbiVersion = MPGOS_B200_ABI_VERSION;
	L.desc = s->desc;
	L.d_TimeDomain = s->var[MPGOS_CVAR_TimeDomain].dev;  L.d_ActualState = s->var[MPGOS_CVAR_ActualState].dev;  L.d_ActualTime = s->var[MPGOS_CVAR_ActualTime].dev;
	L.d_UnitParameters = s->var[MPGOS_CVAR_UnitParameters].dev;  L.d_SystemParameters = s->var[MPGOS_CVAR_SystemParameters].dev;
	L.d_GlobalParameters = s->var[MPGOS_CVAR_GlobalParameters].dev;  L.d_IntegerGlobalParameters = (int*)s->var[MPGOS_CVAR_IntegerGlobalParameters].dev;
	L.d_UnitAccessories = s->var[MPGOS_CVAR_UnitAccessories].dev;  L.d_IntegerUnitAccessories = (int*)s->var[MPGOS_CVAR_IntegerUnitAccessories].dev;
	L.d_SystemAccessories = s->var[MPGOS_CVAR_SystemAccessories].dev;  L.d_IntegerSystemAccessories = (int*)s->var[MPGOS_CVAR_IntegerSystemAccessories].dev;
	L.d_CouplingMatrix = s->var[MPGOS_CVAR_CouplingMatrix].dev;  L.d_CouplingStrength = s->var[MPGOS_CVAR_CouplingStrength].dev;
	L.d_CouplingIndex = (int*)s->var[MPGOS_CVAR_CouplingIndex].dev;  L.d_DenseOutputIndex = (int*)s->var[MPGOS_CVAR_DenseIndex].dev;
	L.d_DenseOutputTimeInstances = s->var[MPGOS_CVAR_DenseTime].dev;  L.d_DenseOutputStates = s->var[MPGOS_CVAR_DenseState].dev;
	L.d_StepsAttempted = s->d_attempted;  L.d_StepsAccepted = s->d_accepted;  L.d_WorkQueue = s->d_queue;
	L.d_LastPhaseSteps = s->d_last_steps;  L.d_DispatchOrder = s->d_order;  L.d_DispatchScratch = s->d_order_scratch;
	L.DispatchHistory = (s->DispatchOrder && s->history_valid) ? s->DispatchOrder : 0;  L.DispatchPilot = (s->DispatchOrder && !s->history_valid) ? s->DispatchOrder : 0;  L.d_SliceState = s->d_slice;
	L.d_SliceParameters = s->d_slice_params;
	L.CouplingMatrixElements = s->var[MPGOS_CVAR_CouplingMatrix].elements;
	L.RelativeTolerance = s->rtol.data();  L.AbsoluteTolerance = s->atol.data();  L.EventTolerance = s->evtol.data();  L.EventDirection = s->evdir.data();
	L.InitialTimeStep = s->InitialTimeStep;  L.MaximumTimeStep = s->MaximumTimeStep;  L.MinimumTimeStep = s->MinimumTimeStep;
	L.TimeStepGrowLimit = s->TimeStepGrowLimit;  L.TimeStepShrinkLimit = s->TimeStepShrinkLimit;  L.DenseOutputMinimumTimeStep = s->DenseOutputMinimumTimeStep;
	L.DenseOutputSaveFrequency = s->DenseOutputSaveFrequency;  L.ActiveSystems = s->ActiveSystems;  L.DeviceMessages = s->DeviceMessages;  L.TensorCoreCoupling = s->TensorCoreCoupling;
	L.MultiProcessorCount = s->sm_count;
}

int mpgos_coupled_solve(mpgos_coupled_solver* s, mpgos_coupled_launcher launcher)
{
	if (!s || !launcher) return CFAIL(MPGOS_ERR_ARGUMENT, "Solve: null solver or launcher");
	for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_solve(c, launcher);  if (rc != MPGOS_OK) return rc; }
	if (!s->shards.empty()) return MPGOS_OK;
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	if (!s->history_valid) { const int rc = ensure_slice_parameters(s);  if (rc != MPGOS_OK) return rc; }
	mpgos_coupled_launch L;
	cfill(s, L);
	const int e = launcher(&L, (void*)s->stream);
	if (e != 0) return CFAIL(MPGOS_ERR_CUDA, "Solve: kernel launch failed -> %s", cudaGetErrorString((cudaError_t)e));
	s->history_valid = true;
	return MPGOS_OK;
}

int mpgos_coupled_solve_timed(mpgos_coupled_solver* s, mpgos_coupled_launcher launcher, int phases, float* elapsed_ms)
{
	if (!s || !launcher || !elapsed_ms || phases < 0) return CFAIL(MPGOS_ERR_ARGUMENT, "solve_timed: bad argument");
	if (!s->shards.empty())
	{
		for (mpgos_coupled_solver* c : s->shards)
		{
			CUDA_TRY(cudaSetDevice(c->desc.Device));
			CUDA_TRY(cudaEventRecord(c->t0, c->stream));
			for (int p = 0; p < phases; p++) { const int rc = mpgos_coupled_solve(c, launcher);  if (rc != MPGOS_OK) return rc; }
			CUDA_TRY(cudaEventRecord(c->t1, c->stream));
		}
		*elapsed_ms = 0;
		for (mpgos_coupled_solver* c : s->shards)
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
	{ const int rc = ensure_slice_parameters(s);  if (rc != MPGOS_OK) return rc; }
	mpgos_coupled_launch L;
	cfill(s, L);
	CUDA_TRY(cudaEventRecord(s->t0, s->stream));
	for (int p = 0; p < phases; p++)
	{
		L.DispatchHistory = (s->DispatchOrder && s->history_valid) ? s->DispatchOrder : 0;  L.DispatchPilot = (s->DispatchOrder && !s->history_valid) ? s->DispatchOrder : 0;  L.d_SliceState = s->d_slice;
		const int e = launcher(&L, (void*)s->stream);
		if (e != 0) return CFAIL(MPGOS_ERR_CUDA, "Solve: kernel launch failed -> %s", cudaGetErrorString((cudaError_t)e));
		s->history_valid = true;
	}
	CUDA_TRY(cudaEventRecord(s->t1, s->stream));
	CUDA_TRY(cudaEventSynchronize(s->t1));
	CUDA_TRY(cudaEventElapsedTime(elapsed_ms, s->t0, s->t1));
	return MPGOS_OK;
}

int mpgos_coupled_insert_sync_point(mpgos_coupled_solver* s)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty()) { for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_insert_sync_point(c);  if (rc != MPGOS_OK) return rc; }  return MPGOS_OK; }
	CUDA_TRY(cudaSetDevice(s->desc.Device)); CUDA_TRY(cudaEventRecord(s->event, s->stream));
	return MPGOS_OK;
}
int mpgos_coupled_sync_solver(mpgos_coupled_solver* s)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty()) { for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_sync_solver(c);  if (rc != MPGOS_OK) return rc; }  return MPGOS_OK; }
	CUDA_TRY(cudaSetDevice(s->desc.Device)); CUDA_TRY(cudaEventSynchronize(s->event));
	return MPGOS_OK;
}
int mpgos_coupled_sync_device(mpgos_coupled_solver* s)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty()) { for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_sync_device(c);  if (rc != MPGOS_OK) return rc; }  return MPGOS_OK; }
	CUDA_TRY(cudaSetDevice(s->desc.Device)); CUDA_TRY(cudaDeviceSynchronize());
	return MPGOS_OK;
}

int mpgos_coupled_read_step_counters(mpgos_coupled_solver* s, unsigned long long* attempted, unsigned long long* accepted)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty())
	{
		std::vector<unsigned long long> a, b;
		for (size_t k = 0; k < s->shards.size(); k++)
		{
			mpgos_coupled_solver* c = s->shards[k];
			a.resize(c->desc.NS);  b.resize(c->desc.NS);
			const int rc = mpgos_coupled_read_step_counters(c, attempted ? a.data() : nullptr, accepted ? b.data() : nullptr);
			if (rc != MPGOS_OK) return rc;
			for (int i = 0; i < c->desc.NS; i++)
			{
				const long long g = cshard_global_id(s->tile, (int)s->shards.size(), (int)k, i);
				if (attempted) attempted[g] = a[i];
				if (accepted)  accepted[g] = b[i];
			}
		}
		return MPGOS_OK;
	}
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	const size_t bytes = 8 * (size_t)s->desc.NS;
	if (attempted) CUDA_TRY(cudaMemcpyAsync(attempted, s->d_attempted, bytes, cudaMemcpyDeviceToHost, s->stream));
	if (accepted)  CUDA_TRY(cudaMemcpyAsync(accepted, s->d_accepted, bytes, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	return MPGOS_OK;
}

int mpgos_coupled_reset_step_counters(mpgos_coupled_solver* s)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty()) { for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_reset_step_counters(c);  if (rc != MPGOS_OK) return rc; }  return MPGOS_OK; }
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	CUDA_TRY(cudaMemsetAsync(s->d_attempted, 0, 8 * (size_t)s->desc.NS, s->stream));
	CUDA_TRY(cudaMemsetAsync(s->d_accepted, 0, 8 * (size_t)s->desc.NS, s->stream));
	return MPGOS_OK;
}

static const int kCState[] = { MPGOS_CVAR_TimeDomain, MPGOS_CVAR_ActualTime, MPGOS_CVAR_ActualState, MPGOS_CVAR_UnitAccessories, MPGOS_CVAR_IntegerUnitAccessories,
                               MPGOS_CVAR_SystemAccessories, MPGOS_CVAR_IntegerSystemAccessories, MPGOS_CVAR_DenseIndex };

int mpgos_coupled_save_state(mpgos_coupled_solver* s)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty()) { for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_save_state(c);  if (rc != MPGOS_OK) return rc; }  return MPGOS_OK; }
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	for (int v : kCState)
	{
		CMirror& m = s->var[v];
		if (!m.bytes) continue;
		if (!m.snap) CUDA_TRY(cudaMalloc(&m.snap, m.bytes));
		CUDA_TRY(cudaMemcpyAsync(m.snap, m.dev, m.bytes, cudaMemcpyDeviceToDevice, s->stream));
	}
	return MPGOS_OK;
}

int mpgos_coupled_restore_state(mpgos_coupled_solver* s)
{
	if (!s) return CFAIL(MPGOS_ERR_ARGUMENT, "null solver");
	if (!s->shards.empty()) { for (mpgos_coupled_solver* c : s->shards) { const int rc = mpgos_coupled_restore_state(c);  if (rc != MPGOS_OK) return rc; }  return MPGOS_OK; }
	CUDA_TRY(cudaSetDevice(s->desc.Device));
	for (int v : kCState)
	{
		CMirror& m = s->var[v];
		if (!m.bytes) continue;
		if (!m.snap) return CFAIL(MPGOS_ERR_ARGUMENT, "restore_state: no snapshot saved");
		CUDA_TRY(cudaMemcpyAsync(m.dev, m.snap, m.bytes, cudaMemcpyDeviceToDevice, s->stream));
	}
	s->history_valid = false;
	return MPGOS_OK;
}

void* mpgos_coupled_stream(mpgos_coupled_solver* s) { return !s ? nullptr : (s->shards.empty() ? (void*)s->stream : (void*)s->shards[0]->stream); }

int mpgos_coupled_builtin_desc(const char* system, int ups, int cbw, int cci, mpgos_coupled_desc* desc)
{
	if (!system || !desc) return CFAIL(MPGOS_ERR_ARGUMENT, "null argument");
	int n = 0;
	const mpgos_b200::CoupledBuiltinEntry* e = mpgos_b200::builtin_entries_coupled(&n);
	for (int i = 0; i < n; i++)
		if (!std::strcmp(e[i].name, system) && e[i].UPS == ups && e[i].CBW == cbw && e[i].CCI == cci)
		{
			desc->UPS = ups; desc->CBW = cbw; desc->CCI = cci; desc->UD = e[i].UD; desc->NC = e[i].NC; desc->NUP = e[i].NUP; desc->NSP = e[i].NSP; desc->NGP = e[i].NGP;
			desc->NiGP = e[i].NiGP; desc->NUA = e[i].NUA; desc->NiUA = e[i].NiUA; desc->NSA = e[i].NSA; desc->NiSA = e[i].NiSA; desc->NE = e[i].NE;
			return MPGOS_OK;
		}
	return CFAIL(MPGOS_ERR_UNKNOWN_SYSTEM, "no built-in coupled system '%s' with UPS=%d CBW=%d CCI=%d", system, ups, cbw, cci);
}

mpgos_coupled_launcher mpgos_coupled_builtin_launcher_p(const char* system, int ups, int cbw, int cci, int algorithm, int dense_output, int precision_bytes)
{
	if (!system) return nullptr;
	int n = 0;
	const mpgos_b200::CoupledBuiltinEntry* e = mpgos_b200::builtin_entries_coupled(&n);
	for (int i = 0; i < n; i++)
		if (!std::strcmp(e[i].name, system) && e[i].UPS == ups && e[i].CBW == cbw && e[i].CCI == cci && e[i].Algorithm == algorithm && e[i].Dense == (dense_output != 0) &&
		    e[i].PrecisionBytes == precision_bytes)
			return e[i].launcher;
	CFAIL(MPGOS_ERR_UNKNOWN_SYSTEM, "no built-in coupled launcher for '%s' UPS=%d CBW=%d CCI=%d algorithm %d dense %d precision %d bytes", system, ups, cbw, cci, algorithm, dense_output, precision_bytes);
	return nullptr;
}

mpgos_coupled_launcher mpgos_coupled_builtin_launcher(const char* system, int ups, int cbw, int cci, int algorithm, int dense_output)
{
	return mpgos_coupled_builtin_launcher_p(system, ups, cbw, cci, algorithm, dense_output, 8);
}

} // extern "C"
