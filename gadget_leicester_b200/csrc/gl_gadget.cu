// gl_gadget.cu - drop-in layer for the reference's AoS arrays: struct particle_data P[] (allvars.h:1131) and
// struct sph_particle_data SphP[] (allvars.h:1384), addressed through the offsetof() table the reference side
// supplies (integration/gl_gadget_glue.c), so the library never needs allvars.h.
//
// The force path itself takes ~18 ms for the 1.6M-particle disc; moving 1.6M x 800 B structs through the host
// caches costs about as much, so this layer is built like a pipeline, not like a wrapper:
//
//   host pool :  pack P[] ->|  pack SphP[]   |  scatter gravity -> P[]  |  scatter density -> P[], SphP[] | scatter hydro
//   PCIe      :   H2D (by super-chunk, behind the pack) ...   D2H gravity |        D2H density            |  D2H hydro
//   GPU       :             |  keys, sort, tree, walk         |  density                  |  hmax, hydro     |
//
// * pack A reads only what tree build + gravity need and all of it comes from P[] (Pos, Vel, Mass, Hsml, OldAcc,
//   Type, TimeBin, Ti_begstep); each 256k-particle super-chunk is sent as soon as its last chunk is packed, so
//   the H2D time hides behind the packing.  The SphP[] fields (VelPred, Entropy, DtEntropy) are packed while
//   the GPU builds the tree and walks it, and reach the device before density needs them.
// * every result block crosses PCIe on a copy stream while the next phase computes, and is scattered by the
//   host pool (all cores but one, dynamic chunks, software prefetch) as soon as it has landed.
// * the only serial tail is "D2H of the hydro results + their scatter into SphP[]".
#include "gl_internal.h"
#include "gl_host.h"
#include <chrono>
#include <sys/mman.h>
#include <algorithm>

#define PACK_CHUNK 8192		// particles per work item of the host pool
#define SUPER_CHUNKS 32		// work items per H2D super-chunk (256k particles, 21 MB)
#define BAD_TYPE_MSG "P[i].Type != 0 inside the gas block [0,N_gas): the reference keeps gas first (SphP[i] belongs to P[i])"
#define PF_DIST 8		// software prefetch distance in structs (measured on the B200 box's Xeon: -25% time)

extern "C" int gl_gadget_bind(gl_ctx *c, const gl_gadget_layout *lay)
{
  if(lay->p_stride <= 0 || lay->sph_stride <= 0) return c->fail("gl_gadget_bind: bad strides");
  c->lay = *lay;
  c->have_lay = true;
  return 0;
}

static int ensure_pinned(gl_ctx *c, void **p, size_t *cap, size_t bytes)
{
  if(bytes <= *cap) return 0;
  if(*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  GL_CUDA(c, cudaMallocHost(p, bytes + bytes / 8));
  *cap = bytes + bytes / 8;
  return 0;
}

// Pull staging (one pinned block, SoA): pos 3n | vel 3n | mass n | hsml n | oldacc n | entropy n | dtentropy n |
// velpred 3n | type n (int) | timebin n (int) | tibegstep n (int)
struct PullStage
{
  double *pos, *vel, *mass, *hsml, *oldacc, *entropy, *dtentropy, *velpred;
  double *sph[5];		// Density, Pressure, DhsmlDensityFactor, DivVel, CurlVel (only pulled with an active list)
  int *type, *timebin, *tibeg;
  static size_t bytes(long long n) { return (size_t) n * (19 * sizeof(double) + 3 * sizeof(int)); }
  PullStage(void *base, long long n)
  {
    pos = (double *) base; vel = pos + 3 * n; mass = vel + 3 * n; hsml = mass + n; oldacc = hsml + n;
    entropy = oldacc + n; dtentropy = entropy + n; velpred = dtentropy + n;
    for(int k = 0; k < 5; k++) sph[k] = velpred + 3 * n + k * n;
    type = (int *) (sph[4] + n); timebin = type + n; tibeg = timebin + n;
  }
};

static int ensure_input_buffers(gl_ctx *c, long long n, long long ngas)
{
  GL_CUDA(c, c->in_pos.ensure(3 * (size_t) n));
  GL_CUDA(c, c->in_vel.ensure(3 * (size_t) n));
  GL_CUDA(c, c->in_mass.ensure((size_t) n));
  GL_CUDA(c, c->in_type.ensure((size_t) n));
  GL_CUDA(c, c->in_hsml.ensure((size_t) n));
  GL_CUDA(c, c->in_oldacc.ensure((size_t) n));
  GL_CUDA(c, c->in_timebin.ensure((size_t) n));
  GL_CUDA(c, c->in_tibegstep.ensure((size_t) n));
  GL_CUDA(c, c->in_entropy.ensure((size_t) n));
  GL_CUDA(c, c->in_dtentropy.ensure((size_t) n));
  GL_CUDA(c, c->in_velpred.ensure(3 * (size_t) std::max<long long>(ngas, 1)));
  if(c->n_tgt_in >= 0)
    {				// active list: SphP state of the inactive particles comes from the host (see submit_pack_B)
      GL_CUDA(c, c->o_density.ensure((size_t) n));
      GL_CUDA(c, c->o_pressure.ensure((size_t) n));
      GL_CUDA(c, c->o_dhsml.ensure((size_t) n));
      GL_CUDA(c, c->o_divvel.ensure((size_t) n));
      GL_CUDA(c, c->o_curlvel.ensure((size_t) n));
    }
  return 0;
}

// ---- pack A: everything the tree build and the gravity walk read, all of it from P[] ----
// Inside the gas block [0,N_gas) Type is 0 by the reference's layout (SphP[i] belongs to P[i], allvars.h:1384; every
// reference loop relies on it), so only the first two cache lines of the struct are touched there (Pos, Vel, Mass,
// OldAcc, Hsml); TimeBin / Ti_begstep, which sit 460 bytes further on, are fetched by pack B, which also verifies
// the Type assumption.  The H2D copies of a super-chunk are enqueued on c->stream by the worker that packs its
// last chunk.
// Particles [0, first) are not packed here: the copy engine fetches them (dma_pull_rows).
static std::shared_ptr<HostJob> submit_pack_A(gl_ctx *c, const char *P, long long first, long long n, long long ngas, PullStage st,
					      std::atomic<int> *sc_left, std::atomic<int> *cuda_err)
{
  const gl_gadget_layout L = c->lay;
  const long long nchunks = (n - first + PACK_CHUNK - 1) / PACK_CHUNK;
  const long long nsuper = (nchunks + SUPER_CHUNKS - 1) / SUPER_CHUNKS;
  for(long long s = 0; s < nsuper; s++)
    sc_left[s].store((int) std::min<long long>(SUPER_CHUNKS, nchunks - s * SUPER_CHUNKS));
  return HostPool::get().submit(n - first, PACK_CHUNK, [=](long long k, long long a, long long b) {
    a += first;
    b += first;
    for(long long i = a; i < b; i++)
      {
	const char *pi = P + i * L.p_stride;
	const bool gas = i < ngas;
	if(i + PF_DIST < b)
	  {
	    const char *pf = pi + PF_DIST * L.p_stride;
	    __builtin_prefetch(pf + L.p_Pos);
	    __builtin_prefetch(pf + L.p_OldAcc);
	    if(!gas) __builtin_prefetch(pf + L.p_Ti_begstep);
	  }
	const double *x = (const double *) (pi + L.p_Pos);
	st.pos[3 * i] = x[0]; st.pos[3 * i + 1] = x[1]; st.pos[3 * i + 2] = x[2];
	const double *v = (const double *) (pi + L.p_Vel);
	st.vel[3 * i] = v[0]; st.vel[3 * i + 1] = v[1]; st.vel[3 * i + 2] = v[2];
	st.mass[i] = *(const double *) (pi + L.p_Mass);
	st.hsml[i] = *(const double *) (pi + L.p_Hsml);
	st.oldacc[i] = *(const double *) (pi + L.p_OldAcc);
	if(gas)
	  st.type[i] = 0;
	else
	  {
	    st.type[i] = *(const short *) (pi + L.p_Type);
	    st.timebin[i] = *(const short *) (pi + L.p_TimeBin);
	    st.tibeg[i] = *(const int *) (pi + L.p_Ti_begstep);
	  }
      }
    const long long s = k / SUPER_CHUNKS;
    if(sc_left[s].fetch_sub(1) == 1)
      {
	// last chunk of this super-chunk: send it
	const long long a0 = first + s * SUPER_CHUNKS * (long long) PACK_CHUNK, b0 = std::min(n, a0 + SUPER_CHUNKS * (long long) PACK_CHUNK), m = b0 - a0;
	cudaSetDevice(c->device);
	cudaError_t e = cudaSuccess;
	auto up = [&](void *dst, const void *src, size_t bytes) {
	  cudaError_t r = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
	  if(r != cudaSuccess) e = r;
	};
	up(c->in_pos.p + 3 * a0, st.pos + 3 * a0, 3 * m * sizeof(double));
	up(c->in_vel.p + 3 * a0, st.vel + 3 * a0, 3 * m * sizeof(double));
	up(c->in_mass.p + a0, st.mass + a0, m * sizeof(double));
	up(c->in_hsml.p + a0, st.hsml + a0, m * sizeof(double));
	up(c->in_oldacc.p + a0, st.oldacc + a0, m * sizeof(double));
	up(c->in_type.p + a0, st.type + a0, m * sizeof(int));
	const long long a1 = std::max(a0, ngas);	// the gas block's time bins come with pack B
	if(a1 < b0)
	  {
	    up(c->in_timebin.p + a1, st.timebin + a1, (b0 - a1) * sizeof(int));
	    up(c->in_tibegstep.p + a1, st.tibeg + a1, (b0 - a1) * sizeof(int));
	  }
	if(e != cudaSuccess) cuda_err->store((int) e);
      }
  });
}

// ---- pack B: the SphP[] fields of the gas block (VelPred, Entropy, DtEntropy) and its TimeBin / Ti_begstep from
// P[]; checks Type == 0 there (bad_type counts violations).  Copies go out on `cs`. ----
// With an active list (with_state) the SphP[] results of EARLIER steps are pulled as well: hydro reads Density,
// Pressure, DhsmlDensityFactor, DivVel, CurlVel of inactive neighbours from memory (hydra.c:1236-1275), and only the
// active particles get new values from this call's density().
static std::shared_ptr<HostJob> submit_pack_B(gl_ctx *c, const char *P, const char *S, long long ngas, PullStage st,
					      std::atomic<int> *left, cudaStream_t cs, cudaEvent_t done_ev,
					      std::atomic<int> *cuda_err, std::atomic<long long> *bad_type, bool with_state)
{
  const gl_gadget_layout L = c->lay;
  const long long nchunks = (ngas + PACK_CHUNK - 1) / PACK_CHUNK;
  left->store((int) nchunks);
  return HostPool::get().submit(ngas, PACK_CHUNK, [=](long long, long long a, long long b) {
    long long bad = 0;
    for(long long i = a; i < b; i++)
      {
	const char *si = S + i * L.sph_stride;
	const char *pi = P + i * L.p_stride;
	if(i + PF_DIST < b)
	  {
	    const char *pf = si + PF_DIST * L.sph_stride;
	    __builtin_prefetch(pf + L.s_Entropy);
	    __builtin_prefetch(pf + L.s_DtEntropy);
	    __builtin_prefetch(pi + PF_DIST * L.p_stride + L.p_Ti_begstep);
	  }
	const double *v = (const double *) (si + L.s_VelPred);
	st.velpred[3 * i] = v[0]; st.velpred[3 * i + 1] = v[1]; st.velpred[3 * i + 2] = v[2];
	st.entropy[i] = *(const double *) (si + L.s_Entropy);
	st.dtentropy[i] = *(const double *) (si + L.s_DtEntropy);
	const int ty = *(const short *) (pi + L.p_Type);
	bad += (ty != 0);
	st.type[i] = ty;
	st.timebin[i] = *(const short *) (pi + L.p_TimeBin);
	st.tibeg[i] = *(const int *) (pi + L.p_Ti_begstep);
	if(with_state)
	  {
	    st.sph[0][i] = *(const double *) (si + L.s_Density);
	    st.sph[1][i] = *(const double *) (si + L.s_Pressure);
	    st.sph[2][i] = *(const double *) (si + L.s_DhsmlDensityFactor);
	    st.sph[3][i] = *(const double *) (si + L.s_DivVel);
	    st.sph[4][i] = *(const double *) (si + L.s_CurlVel);
	  }
      }
    if(bad) bad_type->fetch_add(bad);
    if(left->fetch_sub(1) == 1)
      {
	cudaSetDevice(c->device);
	cudaError_t e = cudaMemcpyAsync(c->in_velpred.p, st.velpred, 3 * ngas * sizeof(double), cudaMemcpyHostToDevice, cs);
	if(e == cudaSuccess) e = cudaMemcpyAsync(c->in_entropy.p, st.entropy, ngas * sizeof(double), cudaMemcpyHostToDevice, cs);
	if(e == cudaSuccess) e = cudaMemcpyAsync(c->in_dtentropy.p, st.dtentropy, ngas * sizeof(double), cudaMemcpyHostToDevice, cs);
	if(e == cudaSuccess) e = cudaMemcpyAsync(c->in_timebin.p, st.timebin, ngas * sizeof(int), cudaMemcpyHostToDevice, cs);
	if(e == cudaSuccess) e = cudaMemcpyAsync(c->in_tibegstep.p, st.tibeg, ngas * sizeof(int), cudaMemcpyHostToDevice, cs);
	if(with_state)
	  {
	    double *dst[5] = { c->o_density.p, c->o_pressure.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p };
	    for(int k = 0; k < 5 && e == cudaSuccess; k++)
	      e = cudaMemcpyAsync(dst[k], st.sph[k], ngas * sizeof(double), cudaMemcpyHostToDevice, cs);
	  }
	if(e == cudaSuccess) e = cudaEventRecord(done_ev, cs);
	if(e != cudaSuccess) cuda_err->store((int) e);
      }
  });
}

// ---------------------------------------------------------------- copy-engine (strided DMA) path ----
// cudaMemcpy2DAsync lets the device side walk the AoS array itself: rows of `width` bytes at the struct stride.  Measured
// on the B200 box (scripts/probes/zerocopy_probe.cu): ~2 ns per row up to ~70 bytes, 36 GB/s for 128-byte rows, no CPU
// involved (hand-written zero-copy kernels reach only 5-20 GB/s).  Pack A keeps all cores busy for ~5 ms with the GPU
// idle, so the two run SIDE BY SIDE: the strided copy fetches the leading 35 % of the gas block, the worker pool packs
// the rest (-1.2 ms end to end).  Needs page-locked host arrays (cudaHostAlloc / cudaHostRegister); anything else falls
// back to the pool alone.  The same trick for the RESULT rows (strided D2H into P[] / SphP[]) was measured and dropped:
// overlapping a running kernel it slowed that kernel by 20-70 %, and in the idle tail of the call it only took the
// host memory bandwidth the pool's scatter needs (25.9 ms with, 25.9 ms without).
static bool dma_usable(const void *P, const void *S)
{
  if(getenv("GL_NO_DMA")) return false;
  cudaPointerAttributes a{}, b{};
  if(cudaPointerGetAttributes(&a, P) != cudaSuccess || cudaPointerGetAttributes(&b, S) != cudaSuccess)
    {
      cudaGetLastError();
      return false;
    }
  return a.type == cudaMemoryTypeHost && b.type == cudaMemoryTypeHost;
}

static double dma_fraction(const char *env, double dflt)
{
  if(const char *e = getenv(env)) return std::min(0.9, std::max(0.0, atof(e)));
  return dflt;
}

struct PullRow
{
  int lo, width, pitch;		// byte range [lo, lo + width) of struct particle_data, device row pitch
  int pos, vel, mass, hsml, oldacc;	// offsets inside the row
};

static bool make_pull_row(const gl_gadget_layout &L, PullRow &R)
{
  const int lo = std::min(std::min(L.p_Pos, L.p_Vel), std::min(std::min(L.p_Mass, L.p_Hsml), L.p_OldAcc));
  const int hi = std::max(std::max(L.p_Pos + 24, L.p_Vel + 24), std::max(std::max(L.p_Mass + 8, L.p_Hsml + 8), L.p_OldAcc + 8));
  R.lo = lo;
  R.width = hi - lo;
  R.pitch = (R.width + 15) & ~15;
  R.pos = L.p_Pos - lo; R.vel = L.p_Vel - lo; R.mass = L.p_Mass - lo; R.hsml = L.p_Hsml - lo; R.oldacc = L.p_OldAcc - lo;
  return R.width <= 256;	// a layout that scatters these fields further apart is left to the pool
}

__global__ void k_rows_to_soa(long long nd, const char *__restrict__ rows, PullRow R, double *__restrict__ pos, double *__restrict__ vel,
			      double *__restrict__ mass, double *__restrict__ hsml, double *__restrict__ oldacc, int *__restrict__ type)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nd) return;
  const char *r = rows + i * R.pitch;
  const double *x = (const double *) (r + R.pos), *v = (const double *) (r + R.vel);
  pos[3 * i] = x[0]; pos[3 * i + 1] = x[1]; pos[3 * i + 2] = x[2];
  vel[3 * i] = v[0]; vel[3 * i + 1] = v[1]; vel[3 * i + 2] = v[2];
  mass[i] = *(const double *) (r + R.mass);
  hsml[i] = *(const double *) (r + R.hsml);
  oldacc[i] = *(const double *) (r + R.oldacc);
  type[i] = 0;			// gas block (pack B verifies it)
}

// particles [0, nd) of the gas block: one strided H2D copy + an unpack kernel, both on c->stream
static int dma_pull_rows(gl_ctx *c, const char *P, long long nd)
{
  PullRow R;
  if(nd <= 0 || !make_pull_row(c->lay, R)) return 0;
  GL_CUDA(c, c->dma_rows.ensure((size_t) nd * R.pitch));
  GL_CUDA(c, cudaMemcpy2DAsync(c->dma_rows.p, R.pitch, P + R.lo, c->lay.p_stride, R.width, nd, cudaMemcpyHostToDevice, c->stream));
  k_rows_to_soa<<<gl_grid(nd, 256), 256, 0, c->stream>>>(nd, c->dma_rows.p, R, c->in_pos.p, c->in_vel.p, c->in_mass.p, c->in_hsml.p,
							 c->in_oldacc.p, c->in_type.p);
  GL_LAUNCH_COUNT(c);
  return 0;
}

static int ensure_copy_streams(gl_ctx *c)
{
  if(!c->copy_stream[0])
    for(int k = 0; k < 2; k++) GL_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream[k], cudaStreamNonBlocking));
  return 0;
}

static void mark_loaded(gl_ctx *c, long long n, long long ngas)
{
  if(c->res_count >= 0 && c->res_first + c->res_count > n)
    {
      c->res_first = 0;
      c->res_count = -1;
    }
  c->n = n;
  gl_fresh_particle_set(c);
  c->g_numpart = n;
  c->g_ngas = ngas;
  c->g_off = 0;
  c->g_type = PullStage(c->pin, n).type;
  c->tree_valid = false;
  c->tree_flag = true;
  c->tgt_dirty = true;
  if(!c->domain_imposed) c->have_domain = false;
}

// the SphP part of the state has reached the device on `cs` (event ev): make c->stream see it and publish VelPred
static int finish_pack_B(gl_ctx *c, long long n, long long ngas, cudaEvent_t ev)
{
  GL_CUDA(c, cudaStreamWaitEvent(c->stream, ev, 0));
  if(n > ngas)
    {				// non-gas particles carry no entropy
      GL_CUDA(c, cudaMemsetAsync(c->in_entropy.p + ngas, 0, (n - ngas) * sizeof(double), c->stream));
      GL_CUDA(c, cudaMemsetAsync(c->in_dtentropy.p + ngas, 0, (n - ngas) * sizeof(double), c->stream));
    }
  c->n_velpred = ngas;
  c->have_velpred = ngas > 0;
  c->tgt_dirty = true;		// the SPH target list depends on the gas block's TimeBin (density_isactive, density.c:1035)
  if(c->tree_valid)
    if(glk_apply_velpred(c)) return 1;
  return 0;
}

extern "C" int gl_gadget_pull(gl_ctx *c, const void *Pv, const void *SphPv, long long NumPart, long long N_gas)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_pull: gl_gadget_bind not called");
  const long long n = NumPart, ngas = std::min(N_gas, NumPart);
  if(n <= 0) return c->fail("gl_gadget_pull: NumPart <= 0");
  if(n >= (1ll << 27) + (1ll << 24)) return c->fail("gl_gadget_pull: n exceeds the 2^27 design limit of the index encoding");
  if(c->let_auto_targets)
    {				// the target list a multi-rank build installed for the previous particle set
      c->n_tgt_in = -1;
      c->tgt_ordered = false;
      c->let_auto_targets = false;
    }
  if(ensure_pinned(c, &c->pin, &c->pin_cap, PullStage::bytes(n))) return 1;
  if(ensure_input_buffers(c, n, ngas)) return 1;
  if(ensure_copy_streams(c)) return 1;
  PullStage st(c->pin, n);
  const long long nsuper = ((n + PACK_CHUNK - 1) / PACK_CHUNK + SUPER_CHUNKS - 1) / SUPER_CHUNKS;
  std::vector<std::atomic<int>> sc_left(nsuper);
  std::atomic<int> b_left{0}, cuda_err{0};
  std::atomic<long long> bad_type{0};
  cudaEventRecord(c->ev[8], c->stream);
  auto ja = submit_pack_A(c, (const char *) Pv, 0, n, ngas, st, sc_left.data(), &cuda_err);
  auto jb = submit_pack_B(c, (const char *) Pv, (const char *) SphPv, ngas, st, &b_left, c->copy_stream[0], c->ev[12], &cuda_err, &bad_type,
			  c->n_tgt_in >= 0);
  ja->wait();
  jb->wait();
  if(cuda_err.load()) return c->fail(std::string("gl_gadget_pull: ") + cudaGetErrorString((cudaError_t) cuda_err.load()));
  if(bad_type.load()) return c->fail(BAD_TYPE_MSG);
  c->have_velpred = false;
  mark_loaded(c, n, ngas);
  if(ngas > 0)
    {
      if(finish_pack_B(c, n, ngas, c->ev[12])) return 1;
    }
  else
    {
      GL_CUDA(c, cudaMemsetAsync(c->in_entropy.p, 0, n * sizeof(double), c->stream));
      GL_CUDA(c, cudaMemsetAsync(c->in_dtentropy.p, 0, n * sizeof(double), c->stream));
    }
  cudaEventRecord(c->ev[9], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]);
  c->st.ms_h2d = ms;
  return 0;
}

// ---------------------------------------------------------------- result scatter ----
// Result staging (pinned): A = gravity (acc 3n | oldacc n | gravcost n floats), B = density (hsml | numngb | density |
// dhsml | divvel | curlvel | pressure | sinkrho | sinkentr | sinkvel 3n), C = hydro (hacc 3n | dtentr n | maxsig n)
struct PushStage
{
  double *A, *B, *C;
  static size_t bytes(long long n) { return (size_t) n * 24 * sizeof(double); }
  PushStage(void *base, long long n)
  {
    A = (double *) base;
    B = A + 5 * n;
    C = B + 13 * n;
  }
};

// the scatter only touches the particles that were evaluated: the active list when one was set (gl_set_targets)
static int target_list(gl_ctx *c, const int **idx, long long *cnt)
{
  if(c->n_tgt_in < 0 || c->let_auto_targets)	// (multi-rank mode lists the own particles by itself: that is "all" here)
    {
      *idx = nullptr;
      *cnt = c->g_numpart;
      return 0;
    }
  if((long long) c->tgt_host.size() != c->n_tgt_in)
    {
      c->tgt_host.resize((size_t) c->n_tgt_in);
      if(c->n_tgt_in > 0)
	{
	  GL_CUDA(c, cudaMemcpyAsync(c->tgt_host.data(), c->tgt_in.p, c->n_tgt_in * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	  GL_CUDA(c, cudaStreamSynchronize(c->stream));
	}
    }
  *idx = c->tgt_host.data();
  *cnt = c->n_tgt_in;
  return 0;
}

// Results are indexed like the global particle list; the host arrays hold [g_off, g_off + g_numpart) of it.
static int d2h_gravity(gl_ctx *c, cudaStream_t s, double *A, long long a = 0)
{
  const long long n = c->g_numpart, o = c->g_off, m = n - a;
  if(m <= 0) return 0;
  GL_CUDA(c, cudaMemcpyAsync(A + 3 * a, c->o_acc.p + 3 * (o + a), 3 * m * sizeof(double), cudaMemcpyDeviceToHost, s));
  GL_CUDA(c, cudaMemcpyAsync(A + 3 * n + a, c->o_oldacc.p + o + a, m * sizeof(double), cudaMemcpyDeviceToHost, s));
  GL_CUDA(c, cudaMemcpyAsync((float *) (A + 4 * n) + a, c->o_gravcost.p + o + a, m * sizeof(float), cudaMemcpyDeviceToHost, s));
  return 0;
}

static int d2h_density(gl_ctx *c, cudaStream_t s, double *B, long long a = 0)
{
  const long long n = c->g_numpart, o = c->g_off, m = n - a;
  if(m <= 0) return 0;
  const double *src[9] = { c->in_hsml.p, c->o_numngb.p, c->o_density.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p,
    c->o_pressure.p, c->o_sinkrho.p, c->o_sinkentr.p };
  const bool have_other = (n > c->g_ngas);	// sink / dust fields only exist for non-gas targets
  for(int k = 0; k < (have_other ? 9 : 7); k++)
    GL_CUDA(c, cudaMemcpyAsync(B + k * n + a, src[k] + o + a, m * sizeof(double), cudaMemcpyDeviceToHost, s));
  if(have_other)
    GL_CUDA(c, cudaMemcpyAsync(B + 9 * n + 3 * a, c->o_sinkvel.p + 3 * (o + a), 3 * m * sizeof(double), cudaMemcpyDeviceToHost, s));
  return 0;
}

// hydro results of the local index range [a,b)
static int d2h_hydro(gl_ctx *c, cudaStream_t s, double *C, long long a, long long b)
{
  const long long n = c->g_numpart, o = c->g_off, m = b - a;
  if(m <= 0) return 0;
  GL_CUDA(c, cudaMemcpyAsync(C + 3 * a, c->o_hacc.p + 3 * (o + a), 3 * m * sizeof(double), cudaMemcpyDeviceToHost, s));
  GL_CUDA(c, cudaMemcpyAsync(C + 3 * n + a, c->o_dtentr.p + o + a, m * sizeof(double), cudaMemcpyDeviceToHost, s));
  GL_CUDA(c, cudaMemcpyAsync(C + 4 * n + a, c->o_maxsig.p + o + a, m * sizeof(double), cudaMemcpyDeviceToHost, s));
  return 0;
}

// reset_swallow: also do density()'s P[i].SwallowID = 0 (density.c:130) here - the field sits right behind
// GravAccel/OldAcc/Hsml/NumNgb, i.e. in a cache line this loop owns anyway, and nothing reads it in between.
static void scatter_gravity(gl_ctx *c, char *P, const double *A, const int *idx, long long cnt, bool reset_swallow, long long first = 0)
{
  const long long n = c->g_numpart;
  const gl_gadget_layout L = c->lay;
  const double *acc = A, *old = A + 3 * n;
  const float *cost = (const float *) (A + 4 * n);
  const int *type = c->g_type;
  HostPool::get().run(cnt - first, PACK_CHUNK, [=](long long, long long a, long long b) {
    a += first;
    b += first;
    for(long long k = a; k < b; k++)
      {
	const long long i = idx ? idx[k] : k;
	char *pi = P + i * L.p_stride;
	if(k + PF_DIST < b)
	  {
	    char *pf = P + (idx ? idx[k + PF_DIST] : k + PF_DIST) * L.p_stride;
	    __builtin_prefetch(pf + L.p_GravAccel, 1);
	    __builtin_prefetch(pf + L.p_GravCost, 1);
	  }
	double *g = (double *) (pi + L.p_GravAccel);
	g[0] = acc[3 * i]; g[1] = acc[3 * i + 1]; g[2] = acc[3 * i + 2];
	*(double *) (pi + L.p_OldAcc) = old[i];
	*(float *) (pi + L.p_GravCost) = cost[i];
	if(reset_swallow)
	  {
	    const int ty = type[i];
	    if(ty == 0 || ty == 5 || ty == 2) *(unsigned int *) (pi + L.p_SwallowID) = 0;
	  }
      }
  });
}

// density results: P[].Hsml, .n.NumNgb, .SwallowID = 0 (density.c:130), sink / dust fields; SphP[].d.Density,
// .h.DhsmlDensityFactor, .v.DivVel, .r.CurlVel, .Pressure (density.c:434-546, :944-988).
// with_rho_p = false leaves SphP[].d.Density / .Pressure to scatter_hydro, which touches their cache lines anyway.
static void scatter_density(gl_ctx *c, char *P, char *S, const double *B, const int *idx, long long cnt, bool with_rho_p, bool reset_swallow,
			    long long first = 0)
{
  const long long n = c->g_numpart, N_gas = c->g_ngas;
  const gl_gadget_layout L = c->lay;
  const double *hs = B, *nn = hs + n, *de = nn + n, *dh = de + n, *dv = dh + n, *cu = dv + n, *pr = cu + n, *sr = pr + n,
    *se = sr + n, *sv = se + n;
  const int *type = c->g_type;	// Type column of the pack staging (still valid)
  const bool have_other = (n > N_gas);
  HostPool::get().run(cnt - first, PACK_CHUNK, [=](long long, long long a, long long b) {
    a += first;
    b += first;
    for(long long k = a; k < b; k++)
      {
	const long long i = idx ? idx[k] : k;
	if(k + PF_DIST < b)
	  {
	    const long long ip = idx ? idx[k + PF_DIST] : k + PF_DIST;
	    char *pf = P + ip * L.p_stride;
	    __builtin_prefetch(pf + L.p_Hsml, 1);
	    if(reset_swallow) __builtin_prefetch(pf + L.p_SwallowID, 1);
	    if(ip < N_gas)
	      {
		char *sf = S + ip * L.sph_stride;
		__builtin_prefetch(sf + L.s_DivVel, 1);
		if(with_rho_p)
		  {
		    __builtin_prefetch(sf + L.s_Pressure, 1);
		    __builtin_prefetch(sf + L.s_Density, 1);
		  }
	      }
	  }
	const int ty = type[i];
	if(!(ty == 0 || ty == 5 || ty == 2)) continue;
	char *pi = P + i * L.p_stride;
	*(double *) (pi + L.p_Hsml) = hs[i];
	*(double *) (pi + L.p_NumNgb) = nn[i];
	if(reset_swallow) *(unsigned int *) (pi + L.p_SwallowID) = 0;
	if(ty == 0 && i < N_gas)
	  {
	    char *si = S + i * L.sph_stride;
	    if(with_rho_p)
	      {
		*(double *) (si + L.s_Pressure) = pr[i];
		*(double *) (si + L.s_Density) = de[i];
	      }
	    *(double *) (si + L.s_DhsmlDensityFactor) = dh[i];
	    *(double *) (si + L.s_DivVel) = dv[i];
	    *(double *) (si + L.s_CurlVel) = cu[i];
	  }
	else if(ty == 5 && have_other)
	  {
	    *(double *) (pi + L.p_SinkDensity) = sr[i];
	    *(double *) (pi + L.p_SinkEntropy) = se[i];
	    double *v = (double *) (pi + L.p_SinkGasVel);
	    v[0] = sv[3 * i]; v[1] = sv[3 * i + 1]; v[2] = sv[3 * i + 2];
	  }
	else if(ty == 2 && have_other)
	  {
	    *(double *) (pi + L.p_DustDensity) = sr[i];
	    *(double *) (pi + L.p_DustEntropy) = se[i];
	    double *v = (double *) (pi + L.p_DustGasVel);
	    v[0] = sv[3 * i]; v[1] = sv[3 * i + 1]; v[2] = sv[3 * i + 2];
	  }
      }
  });
}

// SphP[].a.HydroAccel, .e.DtEntropy, .MaxSignalVel (hydra.c:1931-1940, :583) for list entries [k0,k1); with B != NULL
// also the SphP[].d.Density / .Pressure that scatter_density left behind (same cache lines).
static void scatter_hydro(gl_ctx *c, char *S, const double *C, const double *B, const int *idx, long long k0, long long k1)
{
  const long long n = c->g_numpart, N_gas = c->g_ngas;
  const gl_gadget_layout L = c->lay;
  const double *ha = C, *dt = ha + 3 * n, *ms = dt + n;
  const double *de = B ? B + 2 * n : nullptr, *pr = B ? B + 6 * n : nullptr;
  const int *type = c->g_type;
  HostPool::get().run(k1 - k0, PACK_CHUNK, [=](long long, long long a, long long b) {
    a += k0;
    b += k0;
    for(long long k = a; k < b; k++)
      {
	const long long i = idx ? idx[k] : k;
	if(k + PF_DIST < b)
	  {
	    const long long ip = idx ? idx[k + PF_DIST] : k + PF_DIST;
	    if(ip < N_gas)
	      {
		char *sf = S + ip * L.sph_stride;
		__builtin_prefetch(sf + L.s_MaxSignalVel, 1);
		__builtin_prefetch(sf + L.s_HydroAccel, 1);
	      }
	  }
	if(i >= N_gas || type[i] != 0) continue;
	char *si = S + i * L.sph_stride;
	if(pr)
	  {
	    *(double *) (si + L.s_Pressure) = pr[i];
	    *(double *) (si + L.s_Density) = de[i];
	  }
	*(double *) (si + L.s_MaxSignalVel) = ms[i];
	*(double *) (si + L.s_DtEntropy) = dt[i];
	double *g = (double *) (si + L.s_HydroAccel);
	g[0] = ha[3 * i]; g[1] = ha[3 * i + 1]; g[2] = ha[3 * i + 2];
      }
  });
}

extern "C" int gl_gadget_push(gl_ctx *c, void *Pv, void *SphPv, int what)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_push: gl_gadget_bind not called");
  const long long n = c->g_numpart;
  if(!c->g_type || n <= 0) return c->fail("gl_gadget_push: gl_gadget_pull has not been called for this particle set");
  if(ensure_pinned(c, &c->pin2, &c->pin2_cap, PushStage::bytes(n))) return 1;
  PushStage st(c->pin2, n);
  const int *idx;
  long long cnt;
  if(target_list(c, &idx, &cnt)) return 1;
  if(what & 1)
    {
      if(d2h_gravity(c, c->stream, st.A)) return 1;
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      scatter_gravity(c, (char *) Pv, st.A, idx, cnt, false);
    }
  if(what & 2)
    {
      if(d2h_density(c, c->stream, st.B)) return 1;
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      scatter_density(c, (char *) Pv, (char *) SphPv, st.B, idx, cnt, true, true);
    }
  if(what & 4)
    {
      if(d2h_hydro(c, c->stream, st.C, 0, n)) return 1;
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      scatter_hydro(c, (char *) SphPv, st.C, nullptr, idx, 0, cnt);
    }
  return 0;
}

// ---------------------------------------------------------------- the pipelined drop-in call ----
namespace
{
// Wait for a copy to land WITHOUT parking inside the CUDA runtime: a thread blocked in cudaEventSynchronize was
// measured to stall cudaStreamSynchronize calls of the GPU-driving thread until the event fired (2.7 ms per
// step); polling with short sleeps costs no core and holds no driver lock.
inline cudaError_t wait_event_polling(cudaEvent_t ev)
{
  cudaError_t e;
  while((e = cudaEventQuery(ev)) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(40));
  return e;
}

// hand-over between the thread that drives the GPU and the thread that scatters results
struct Stages
{
  std::mutex m;
  std::condition_variable cv;
  int posted = 0;		// number of result blocks whose D2H has been enqueued (1 gravity, 2 density, 3 hydro)
  bool abort = false;
  void post(int k)
  {
    {
      std::lock_guard<std::mutex> lk(m);
      posted = k;
    }
    cv.notify_all();
  }
  void stop()
  {
    {
      std::lock_guard<std::mutex> lk(m);
      abort = true;
    }
    cv.notify_all();
  }
  bool wait_for(int k)
  {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return abort || posted >= k; });
    return posted >= k;
  }
};
}

// compute_accelerations (accel.c:27, force part :61-106) on the reference's host arrays.
extern "C" int gl_gadget_compute_accelerations(gl_ctx *c, void *Pv, void *SphPv, long long NumPart, long long N_gas, int mode)
{
  if(c->comm && c->comm->world > 1)
    {
      // multi-rank mode: this rank's own P[] / SphP[]; the phases are collective and may be repeated (gl_let.cu), so
      // they run through the ordinary entry points, without the software pipeline of the single-rank path below
      // (collective, repeatable entry points); every result block goes home behind the next phase's kernel launch and
      // is scattered by a host thread while that phase computes (gl_gadget_push_segment_async)
      cudaSetDevice(c->device);
      (void) mode;
      if(gl_gadget_pull(c, Pv, SphPv, NumPart, N_gas)) return 1;
      gl_set_tree_reconstruct_flag(c, 1);
      const bool first = (c->par.TypeOfOpeningCriterion == 1 && c->par.Ti_Current == 0 && c->par.ErrTolTheta != 0);
      if(gl_gravity_tree(c)) return 1;	// accel.c:61
      if(first)
	if(gl_gravity_tree(c)) return 1;	// accel.c:63-67
      if(gl_gadget_push_segment_async(c, Pv, SphPv, 1)) return 1;
      if(gl_prepare_targets(c)) return 1;
      if(c->n_tgt_sph > 0)
	{
	  if(gl_density(c)) return 1;	// accel.c:84
	  if(gl_gadget_push_segment_async(c, Pv, SphPv, 2)) return 1;
	  if(gl_force_update_hmax(c)) return 1;	// accel.c:96
	  if(gl_hydro_force(c)) return 1;	// accel.c:106
	  if(gl_gadget_push_segment_async(c, Pv, SphPv, 4)) return 1;
	}
      return gl_gadget_push_wait(c);
    }
  (void) mode;
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_compute_accelerations: gl_gadget_bind not called");
  const bool trace = getenv("GL_DEBUG_E2E") != nullptr;
  auto now_ms = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now_ms();
  double t_packA = 0, t_grav = 0, t_dens = 0, t_hyd = 0, t_sg = 0, t_sd = 0;
  const long long n = NumPart, ngas = std::min(N_gas, NumPart);
  if(n <= 0) return c->fail("gl_gadget_compute_accelerations: NumPart <= 0");
  if(n >= (1ll << 27) + (1ll << 24)) return c->fail("gl_gadget_compute_accelerations: n exceeds the 2^27 design limit");
  if(ensure_pinned(c, &c->pin, &c->pin_cap, PullStage::bytes(n))) return 1;
  if(ensure_pinned(c, &c->pin2, &c->pin2_cap, PushStage::bytes(n))) return 1;
  if(ensure_input_buffers(c, n, ngas)) return 1;
  if(ensure_copy_streams(c)) return 1;
  cudaStream_t cs0 = c->copy_stream[0], cs1 = c->copy_stream[1];
  PullStage in(c->pin, n);
  PushStage out(c->pin2, n);
  char *P = (char *) Pv, *S = (char *) SphPv;

  // ---- pull: pack A (H2D behind it on c->stream), then pack B in the background ----
  const long long nsuper = ((n + PACK_CHUNK - 1) / PACK_CHUNK + SUPER_CHUNKS - 1) / SUPER_CHUNKS;
  std::vector<std::atomic<int>> sc_left(nsuper);
  std::atomic<int> b_left{0}, cuda_err{0};
  std::atomic<long long> bad_type{0};
  // the copy engine fetches the leading part of the gas block while the pool packs the rest
  const bool dma = dma_usable(P, S) && c->n_tgt_in < 0;
  PullRow prow;
  const long long nd_pull = (dma && make_pull_row(c->lay, prow)) ? (long long) (dma_fraction("GL_DMA_PULL", 0.35) * ngas) : 0;
  if(dma_pull_rows(c, P, nd_pull)) return 1;
  auto ja = submit_pack_A(c, P, nd_pull, n, ngas, in, sc_left.data(), &cuda_err);
  auto jb = submit_pack_B(c, P, S, ngas, in, &b_left, cs0, c->ev[12], &cuda_err, &bad_type, c->n_tgt_in >= 0);
  ja->wait();
  t_packA = now_ms();
  c->have_velpred = false;
  mark_loaded(c, n, ngas);

  const int *idx;
  long long cnt;
  int rc = 0;
  Stages stages;
  std::thread scatterer;
  const bool saved_blocking = c->blocking_sync;
  c->blocking_sync = true;
  auto bail = [&](int code) {
    c->post_launch = nullptr;
    stages.stop();
    jb->wait();
    if(scatterer.joinable()) scatterer.join();
    c->blocking_sync = saved_blocking;
    return code;
  };
#define GLG_CUDA(call)                                                                                         \
  do {                                                                                                         \
    cudaError_t e__ = (call);                                                                                  \
    if(e__ != cudaSuccess) return bail(c->fail(std::string(#call) + ": " + cudaGetErrorString(e__)));         \
  } while(0)
  if(cuda_err.load()) return bail(c->fail(std::string("pack/H2D: ") + cudaGetErrorString((cudaError_t) cuda_err.load())));
  if(target_list(c, &idx, &cnt)) return bail(1);
  if(ngas == 0)
    {
      GLG_CUDA(cudaMemsetAsync(c->in_entropy.p, 0, n * sizeof(double), c->stream));
      GLG_CUDA(cudaMemsetAsync(c->in_dtentropy.p, 0, n * sizeof(double), c->stream));
    }

  // ---- the scatter thread: waits for each result block to land, then runs the host pool on it ----
  // blocks: 1 gravity, 2 density, 3 hydro of P[] indices [0,split), 4 hydro of [split,n)
  cudaEvent_t evG = c->ev_res[0], evD = c->ev_res[1], evH1 = c->ev_res[2], evH2 = c->ev_res[3];
  const long long split = idx ? cnt : ngas / 2;	// an explicit active list is not split
  scatterer = std::thread([&]() {
    cudaSetDevice(c->device);
    jb->wait();			// FIFO pool: the SphP pack is ahead of every scatter anyway
    if(!stages.wait_for(1)) return;
    wait_event_polling(evG);
    scatter_gravity(c, P, out.A, idx, cnt, true);
    t_sg = now_ms();
    if(!stages.wait_for(2)) return;
    wait_event_polling(evD);
    scatter_density(c, P, S, out.B, idx, cnt, false, false);
    t_sd = now_ms();
    if(!stages.wait_for(3)) return;
    wait_event_polling(evH1);
    scatter_hydro(c, S, out.C, out.B, idx, 0, split);
    if(!stages.wait_for(4)) return;
    wait_event_polling(evH2);
    scatter_hydro(c, S, out.C, out.B, idx, split, cnt);
  });
  int hook_rc = 0;

  // ---- GPU: gravity (tree build inside) ----
  const bool first = (c->par.TypeOfOpeningCriterion == 1 && c->par.Ti_Current == 0 && c->par.ErrTolTheta != 0);
  c->tgt_dirty = true;
  if(glg_gravity(c)) return bail(1);	// accel.c:61
  if(first)
    if(glg_gravity(c)) return bail(1);	// accel.c:63-67
  t_grav = now_ms();
  GLG_CUDA(cudaEventRecord(c->ev[11], c->stream));
  // the gravity block goes out once the density kernel has been launched (see gl_ctx::post_launch)
  c->post_launch = [&]() {
    if(cudaStreamWaitEvent(cs1, c->ev[11], 0) != cudaSuccess || d2h_gravity(c, cs1, out.A) || cudaEventRecord(evG, cs1) != cudaSuccess)
      {
	hook_rc = 1;
	stages.stop();		// nothing valid will land: the scatter thread must not touch P[]
	return;
      }
    stages.post(1);
  };

  // ---- the SphP part of the state must be on the device now ----
  jb->wait();
  if(cuda_err.load()) return bail(c->fail(std::string("pack/H2D: ") + cudaGetErrorString((cudaError_t) cuda_err.load())));
  if(bad_type.load()) return bail(c->fail(BAD_TYPE_MSG));
  if(ngas > 0)
    if(finish_pack_B(c, n, ngas, c->ev[12])) return bail(1);

  if(gl_prepare_targets(c)) return bail(1);
  const bool sph = c->n_tgt_sph > 0;
  if(sph)
    {
      rc = gls_density(c);	// accel.c:84 (fires the gravity hook behind its kernel launch)
      t_dens = now_ms();
      if(rc || hook_rc) return bail(1);
      GLG_CUDA(cudaEventRecord(c->ev[13], c->stream));
      rc = glt_update_hmax(c);	// accel.c:96
      if(rc) return bail(rc);
      // accel.c:106 as two launches: [0,split) first, so its results travel and are scattered while the rest runs
      long long n_lo = c->n_tgt_gas;
      const int *tg = c->tgt_gas.p;
      if(gls_hydro_begin(c)) return bail(1);
      if(!idx && c->n_tgt_gas > 0)
	{
	  if(gl_split_gas_targets(c, split, &n_lo)) return bail(1);
	  tg = c->tgt_gas_part.p;
	}
      if(gls_hydro_launch(c, tg, n_lo)) return bail(1);
      GLG_CUDA(cudaEventRecord(c->ev[14], c->stream));
      if(gls_hydro_launch(c, tg + n_lo, c->n_tgt_gas - n_lo)) return bail(1);
      c->post_launch = [&]() {
	if(cudaStreamWaitEvent(cs0, c->ev[13], 0) != cudaSuccess || d2h_density(c, cs0, out.B) || cudaEventRecord(evD, cs0) != cudaSuccess)
	  {
	    hook_rc = 1;
	    stages.stop();
	    return;
	  }
	stages.post(2);
	if(cudaStreamWaitEvent(cs1, c->ev[14], 0) != cudaSuccess || d2h_hydro(c, cs1, out.C, 0, idx ? n : split) ||
	   cudaEventRecord(evH1, cs1) != cudaSuccess)
	  {
	    hook_rc = 1;
	    stages.stop();
	    return;
	  }
	stages.post(3);
      };
      rc = gls_hydro_end(c);	// fires the hook, then waits for both launches
      t_hyd = now_ms();
      if(rc || hook_rc) return bail(1);
      if(!idx)
	if(d2h_hydro(c, c->stream, out.C, split, n)) return bail(1);
      GLG_CUDA(cudaEventRecord(evH2, c->stream));
      stages.post(4);
    }
  else
    {
      c->fire_post_launch();	// gravity-only particle set: the scatter thread ends after the gravity block
      if(hook_rc) return bail(1);
      stages.stop();
    }
  scatterer.join();
  c->blocking_sync = saved_blocking;
  if(trace)
    fprintf(stderr, "[gl] e2e ms since entry: pack A %.1f | gravity done %.1f | gravity scattered %.1f | density done %.1f | "
	    "density scattered %.1f | hydro done %.1f | exit %.1f  (%d pool threads)\n",
	    t_packA - t_start, t_grav - t_start, t_sg - t_start, t_dens - t_start, t_sd - t_start, t_hyd - t_start,
	    now_ms() - t_start, HostPool::get().threads());
  if(trace)
    fprintf(stderr, "[gl]     kernel events ms: keys+sort %.2f build %.2f moments %.2f walk %.2f density %.2f hmax %.2f hydro %.2f\n",
	    c->st.ms_keys_sort, c->st.ms_treebuild, c->st.ms_moments, c->st.ms_gravity_walk, c->st.ms_density, c->st.ms_hmax, c->st.ms_hydro);
  return rc;
#undef GLG_CUDA
}

// ---------------------------------------------------------------- multi-GPU: one rank's P[] / SphP[] ----
// With the sources replicated by all-gathers (gl_load_particles_gathered) every rank still owns its own host arrays,
// like every MPI rank of the reference does.  pack_segment turns them into this rank's all-gather block;
// push_segment_async returns the rank's slice of a result block to them while the next phase computes.
extern "C" int gl_gadget_pack_segment(gl_ctx *c, const void *Pv, const void *SphPv, long long NumPart, long long N_gas, long long m,
				       long long global_offset, double *f64_dev, int *i32_dev)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_pack_segment: gl_gadget_bind not called");
  const long long n = NumPart, ngas = std::min(N_gas, NumPart);
  if(n < 0 || n > m) return c->fail("gl_gadget_pack_segment: segment longer than the block (m)");
  const size_t fbytes = (size_t) m * 11 * sizeof(double), ibytes = (size_t) m * 3 * sizeof(int);
  if(ensure_pinned(c, &c->pin, &c->pin_cap, fbytes + ibytes)) return 1;
  double *f = (double *) c->pin;
  int *q = (int *) (f + 11 * m);
  const gl_gadget_layout L = c->lay;
  const char *P = (const char *) Pv, *S = (const char *) SphPv;
  HostPool::get().run(n, PACK_CHUNK, [=](long long, long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	const char *pi = P + i * L.p_stride;
	const bool gas = i < ngas;
	const char *si = S + i * L.sph_stride;
	if(i + PF_DIST < b)
	  {
	    const char *pf = pi + PF_DIST * L.p_stride;
	    __builtin_prefetch(pf + L.p_Pos);
	    __builtin_prefetch(pf + L.p_OldAcc);
	    __builtin_prefetch(pf + L.p_Ti_begstep);
	    if(gas)
	      {
		__builtin_prefetch(si + PF_DIST * L.sph_stride + L.s_Entropy);
		__builtin_prefetch(si + PF_DIST * L.sph_stride + L.s_DtEntropy);
	      }
	  }
	const double *x = (const double *) (pi + L.p_Pos);
	f[3 * i] = x[0]; f[3 * i + 1] = x[1]; f[3 * i + 2] = x[2];
	f[3 * m + i] = *(const double *) (pi + L.p_Mass);
	f[4 * m + i] = *(const double *) (pi + L.p_Hsml);
	f[5 * m + i] = *(const double *) (pi + L.p_OldAcc);
	const double *v = gas ? (const double *) (si + L.s_VelPred) : (const double *) (pi + L.p_Vel);
	f[6 * m + 3 * i] = v[0]; f[6 * m + 3 * i + 1] = v[1]; f[6 * m + 3 * i + 2] = v[2];
	f[9 * m + i] = gas ? *(const double *) (si + L.s_Entropy) : 0.0;
	f[10 * m + i] = gas ? *(const double *) (si + L.s_DtEntropy) : 0.0;
	q[i] = *(const short *) (pi + L.p_Type);
	q[m + i] = *(const short *) (pi + L.p_TimeBin);
	q[2 * m + i] = *(const int *) (pi + L.p_Ti_begstep);
      }
  });
  GL_CUDA(c, cudaMemcpyAsync(f64_dev, f, fbytes, cudaMemcpyHostToDevice, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(i32_dev, q, ibytes, cudaMemcpyHostToDevice, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->g_numpart = n;
  c->g_ngas = ngas;
  c->g_off = global_offset;
  c->g_type = q;
  return 0;
}

namespace
{
struct AsyncPush
{
  std::mutex m;
  std::condition_variable cv;
  bool enqueued = false, failed = false;
};
}

// what = 1 gravity, 2 density or 4 hydro (one block per call).  The D2H of the block is enqueued behind the next
// phase's kernel launch (gl_ctx::post_launch), a thread scatters it when it has landed; gl_gadget_push_wait joins.
extern "C" int gl_gadget_push_segment_async(gl_ctx *c, void *Pv, void *SphPv, int what)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_push_segment_async: gl_gadget_bind not called");
  if(what != 1 && what != 2 && what != 4) return c->fail("gl_gadget_push_segment_async: what must be 1, 2 or 4");
  const long long n = c->g_numpart;
  if(!c->g_type || n <= 0) return c->fail("gl_gadget_push_segment_async: no host arrays packed");
  if(ensure_pinned(c, &c->pin2, &c->pin2_cap, PushStage::bytes(n))) return 1;
  if(ensure_copy_streams(c)) return 1;
  const int slot = what == 1 ? 0 : (what == 2 ? 1 : 2);
  PushStage st(c->pin2, n);
  cudaEvent_t ready = c->ev[11 + slot], landed = c->ev_res[slot];
  cudaStream_t cs = c->copy_stream[slot & 1];
  c->fire_post_launch();	// a block still waiting for "the next launch" goes out now
  GL_CUDA(c, cudaEventRecord(ready, c->stream));
  auto ap = std::make_shared<AsyncPush>();
  char *P = (char *) Pv, *S = (char *) SphPv;
  c->post_launch = [=]() {
    bool bad = cudaStreamWaitEvent(cs, ready, 0) != cudaSuccess;
    if(!bad) bad = (what == 1) ? d2h_gravity(c, cs, st.A) : (what == 2) ? d2h_density(c, cs, st.B) : d2h_hydro(c, cs, st.C, 0, n);
    if(!bad) bad = cudaEventRecord(landed, cs) != cudaSuccess;
    {
      std::lock_guard<std::mutex> lk(ap->m);
      ap->enqueued = true;
      ap->failed = bad;
    }
    ap->cv.notify_all();
  };
  c->push_threads.emplace_back([=]() {
    cudaSetDevice(c->device);
    {
      std::unique_lock<std::mutex> lk(ap->m);
      ap->cv.wait(lk, [&] { return ap->enqueued; });
    }
    if(ap->failed || wait_event_polling(landed) != cudaSuccess)
      {
	c->push_failed = 1;
	return;
      }
    if(what == 1)
      scatter_gravity(c, P, st.A, nullptr, n, false);
    else if(what == 2)
      scatter_density(c, P, S, st.B, nullptr, n, true, true);
    else
      scatter_hydro(c, S, st.C, nullptr, nullptr, 0, n);
  });
  return 0;
}

extern "C" int gl_gadget_push_wait(gl_ctx *c)
{
  cudaSetDevice(c->device);
  c->fire_post_launch();
  for(auto &t : c->push_threads)
    if(t.joinable()) t.join();
  c->push_threads.clear();
  if(c->push_failed)
    {
      c->push_failed = 0;
      return c->fail("gl_gadget_push_segment_async: a result block could not be copied back");
    }
  return 0;
}

// ---------------------------------------------------------------- lazy drift of the sources (host) ----
static int drift_particles(const gl_gadget_layout *lay, void *Pv, void *SphPv, const int *list, long long nlist, long long N_gas,
			   int time1, double Timebase_interval, double MinGasHsml, long long *ndrifted)
{
  const gl_gadget_layout L = *lay;
  char *P = (char *) Pv, *S = (char *) SphPv;
  std::atomic<long long> moved{0}, ahead{0};
  HostPool::get().run(nlist, PACK_CHUNK, [&, L, P, S](long long, long long a, long long b) {
    long long m = 0, bad = 0;
    for(long long k = a; k < b; k++)
      {
	const long long i = list ? list[k] : k;
	char *pi = P + i * L.p_stride;
	const int type = *(const short *) (pi + L.p_Type);
	if(type == 3) continue;	// predict.c:134 (-DVIRTUAL)
	int *tic = (int *) (pi + L.p_Ti_current);
	const int time0 = *tic;
	if(time1 < time0)
	  {
	    bad++;		// predict.c:148-152 endrun(12)
	    continue;
	  }
	if(time1 == time0) continue;
	// predict.c:165: dt_drift = dt_gravkick = dt_hydrokick = (time1 - time0) * All.Timebase_interval
	const double dt = (time1 - time0) * Timebase_interval;
	double *pos = (double *) (pi + L.p_Pos);
	const double *vel = (const double *) (pi + L.p_Vel);
	for(int j = 0; j < 3; j++) pos[j] += vel[j] * dt;	// predict.c:169-170
	if(type == 0 && i < N_gas)
	  {
	    char *si = S + i * L.sph_stride;
	    double *vp = (double *) (si + L.s_VelPred);
	    const double *ga = (const double *) (pi + L.p_GravAccel), *ha = (const double *) (si + L.s_HydroAccel);
	    for(int j = 0; j < 3; j++) vp[j] += ga[j] * dt + ha[j] * dt;	// predict.c:185-186
	    if(L.s_DragAccel >= 0)
	      {
		const double *da = (const double *) (si + L.s_DragAccel);
		for(int j = 0; j < 3; j++) vp[j] += da[j] * dt;	// predict.c:194-196 (-DDUST)
	      }
	    const double divv = *(const double *) (si + L.s_DivVel);
	    double *rho = (double *) (si + L.s_Density), *h = (double *) (pi + L.p_Hsml);
	    *rho *= exp(-divv * dt);	// predict.c:204
	    *h *= exp(0.333333333333 * divv * dt);	// predict.c:205
	    if(*h < MinGasHsml) *h = MinGasHsml;
	    const int tb = *(const short *) (pi + L.p_TimeBin);
	    const int dt_step = tb ? (1 << tb) : 0;
	    const double dt_entr = (time1 - (*(const int *) (pi + L.p_Ti_begstep) + dt_step / 2)) * Timebase_interval;	// predict.c:213
	    // predict.c:217 (neither MHM nor SOFTEREQS in the shipped build)
	    *(double *) (si + L.s_Pressure) =
	      (*(const double *) (si + L.s_Entropy) + *(const double *) (si + L.s_DtEntropy) * dt_entr) * pow(*rho, 7.0 / 5);
	  }
	*tic = time1;
	m++;
      }
    if(m) moved.fetch_add(m);
    if(bad) ahead.fetch_add(bad);
  });
  if(ndrifted) *ndrifted = moved.load();
  return ahead.load() ? 1 : 0;
}

extern "C" int gl_gadget_drift(const gl_gadget_layout *lay, void *P, void *SphP, long long NumPart, long long N_gas, int time1,
			       double Timebase_interval, double MinGasHsml, long long *ndrifted)
{
  return drift_particles(lay, P, SphP, nullptr, NumPart, N_gas, time1, Timebase_interval, MinGasHsml, ndrifted);
}

// the drift of the ACTIVE particles at a synchronisation point (find_next_sync_point_and_drift, run.c:325-329: "drift the
// active particles, others will be drifted on the fly if needed"): only the listed particles are touched
extern "C" int gl_gadget_drift_list(const gl_gadget_layout *lay, void *P, void *SphP, const int *list, long long nlist, long long N_gas,
				    int time1, double Timebase_interval, double MinGasHsml, long long *ndrifted)
{
  if(nlist > 0 && !list) return 1;
  return drift_particles(lay, P, SphP, list, nlist, N_gas, time1, Timebase_interval, MinGasHsml, ndrifted);
}

// ---------------------------------------------------------------- reorder_gas / reorder_particles (host) ----
extern "C" int gl_gadget_reorder(const gl_gadget_layout *lay, void *Pv, void *SphPv, long long NumPart, long long N_gas, const int *order)
{
  const long long ps = lay->p_stride, ss = lay->sph_stride;
  if(NumPart <= 0) return 0;
  for(long long k = 0; k < NumPart; k++)	// a gas slot must take a gas particle (the two blocks are ordered separately)
    if(order[k] < 0 || order[k] >= NumPart || ((k < N_gas) != (order[k] < N_gas))) return 1;
  char *P = (char *) Pv, *S = (char *) SphPv;
  // scratch of the size of the arrays: ask for huge pages, first-touch faults of 4 KB pages cost more than the copies
  auto big_alloc = [](size_t bytes) -> char * {
    const size_t two_mb = (size_t) 2 << 20;
    void *p = nullptr;
    if(posix_memalign(&p, two_mb, (bytes + two_mb - 1) / two_mb * two_mb) != 0) return nullptr;
    madvise(p, (bytes + two_mb - 1) / two_mb * two_mb, MADV_HUGEPAGE);
    return (char *) p;
  };
  char *tp = big_alloc((size_t) NumPart * ps), *tsph = N_gas > 0 ? big_alloc((size_t) N_gas * ss) : nullptr;
  if(!tp || (N_gas > 0 && !tsph))
    {
      free(tp);
      free(tsph);
      return 1;
    }
  HostPool &pool = HostPool::get();
  pool.run(NumPart, PACK_CHUNK, [=](long long, long long a, long long b) {
    memcpy(tp + a * ps, P + a * ps, (size_t) (b - a) * ps);
    if(a < N_gas) memcpy(tsph + a * ss, S + a * ss, (size_t) (std::min(b, N_gas) - a) * ss);
  });
  pool.run(NumPart, PACK_CHUNK, [=](long long, long long a, long long b) {
    for(long long k = a; k < b; k++)
      {
	memcpy(P + k * ps, tp + (long long) order[k] * ps, (size_t) ps);
	if(k < N_gas) memcpy(S + k * ss, tsph + (long long) order[k] * ss, (size_t) ss);
      }
  });
  free(tp);
  free(tsph);
  return 0;
}

// ---------------------------------------------------------------- peano_hilbert_order on the host ----
// peano.c:25-92 with the keys the domain decomposition has already computed (the reference's global Key[],
// domain.c:1750): (key, index) records of the gas block and of the rest are sorted separately - the reference's merge
// sort takes the left element on equal keys (peano.c:551), i.e. it is stable, and so is this one - and the structs are
// permuted accordingly.  The sort runs on all host cores: contiguous stretches are sorted independently, then merged
// pairwise, level by level.
namespace
{
struct KeyIdx
{
  unsigned long long key;
  int index;
};

void sort_block(const unsigned long long *Key, long long first, long long n, int *order)
{
  if(n <= 0) return;
  std::vector<KeyIdx> a((size_t) n), b((size_t) n);
  HostPool &pool = HostPool::get();
  const long long nth = std::max(1, pool.threads());
  long long stretch = (n + nth - 1) / nth;
  if(stretch < 4096) stretch = 4096;
  const long long nst = (n + stretch - 1) / stretch;
  pool.run(nst, 1, [&](long long, long long s0, long long s1) {
    for(long long s = s0; s < s1; s++)
      {
	const long long lo = s * stretch, hi = std::min(n, lo + stretch);
	for(long long i = lo; i < hi; i++) a[(size_t) i] = KeyIdx{Key[first + i], (int) (first + i)};
	std::stable_sort(a.begin() + lo, a.begin() + hi, [](const KeyIdx &x, const KeyIdx &y) { return x.key < y.key; });
      }
  });
  KeyIdx *src = a.data(), *dst = b.data();
  for(long long width = stretch; width < n; width *= 2)
    {
      const long long npairs = (n + 2 * width - 1) / (2 * width);
      pool.run(npairs, 1, [&](long long, long long p0, long long p1) {
	for(long long p = p0; p < p1; p++)
	  {
	    const long long lo = p * 2 * width, mid = std::min(n, lo + width), hi = std::min(n, lo + 2 * width);
	    // std::merge keeps the element of the first range on equal keys: the reference's "b1->key <= b2->key"
	    std::merge(src + lo, src + mid, src + mid, src + hi, dst + lo, [](const KeyIdx &x, const KeyIdx &y) { return x.key < y.key; });
	  }
      });
      std::swap(src, dst);
    }
  for(long long i = 0; i < n; i++) order[first + i] = src[i].index;
}
}				// namespace

extern "C" int gl_gadget_peano_hilbert_order(const gl_gadget_layout *lay, void *P, void *SphP, long long NumPart, long long N_gas,
					     const unsigned long long *Key, int *order_out)
{
  if(NumPart <= 0) return 0;
  if(N_gas < 0 || N_gas > NumPart || NumPart > 0x7fffffffLL) return 1;
  std::vector<int> order((size_t) NumPart);
  const auto t0 = std::chrono::steady_clock::now();
  sort_block(Key, 0, N_gas, order.data());
  sort_block(Key, N_gas, NumPart - N_gas, order.data());
  const auto t1 = std::chrono::steady_clock::now();
  if(order_out) memcpy(order_out, order.data(), sizeof(int) * (size_t) NumPart);
  const int rc = gl_gadget_reorder(lay, P, SphP, NumPart, N_gas, order.data());
  if(getenv("GL_DEBUG_E2E"))
    fprintf(stderr, "peano_hilbert_order: sort %.2f ms, reorder %.2f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
	    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  return rc;
}

// ---------------------------------------------------------------- the top-level tree of the domain decomposition ----
// domain_determineTopTree (domain.c:1738-1916) for one rank: the string of Peano-Hilbert keys is chopped into eight
// segments recursively while a segment holds more than 1/(TOPNODEFACTOR * MULTIPLEDOMAINS * NTask) of the particles or
// of the cost, or more than 80 % of its parent's (domain_check_for_local_refine, domain.c:1524-1590).  The reference
// sorts (key, index) records serially and then walks every refined node's particle range once per level; here the sort
// runs on all host cores and the ranges are found by bisection in the sorted keys.  What must stay exactly as in the
// reference is the ARITHMETIC of the decisions: the root cost is the sum over the particles in P[] order, a daughter's
// cost the sum over its range in sorted order, each started from zero (FP64 addition is not associative), and the
// node numbering, which follows the depth-first order of the recursion.
namespace
{
struct TopBuild
{
  gl_topnode *T;
  int n, max_nodes;
  const unsigned long long *key;	// sorted
  const double *cost;		// in the same order
  double countlimit, costlimit;

  int refine(int i)
  {
    int flag = 0;
    if(T[i].Parent >= 0)
      if(T[i].Count > 0.8 * T[T[i].Parent].Count || T[i].Cost > 0.8 * T[T[i].Parent].Cost) flag = 1;
    if((T[i].Count > countlimit || T[i].Cost > costlimit || flag == 1) && T[i].Size >= 8)
      {
	if(n + 8 > max_nodes) return 1;
	T[i].Daughter = n;
	for(int j = 0; j < 8; j++)
	  {
	    gl_topnode &d = T[n + j];
	    d.Daughter = -1;
	    d.Parent = i;
	    d.Size = T[i].Size >> 3;
	    d.StartKey = T[i].StartKey + j * d.Size;
	    d.PIndex = T[i].PIndex;
	    d.Count = 0;
	    d.Cost = 0;
	    d.Leaf = 0;
	  }
	n += 8;
	const int first = T[i].Daughter;
	const long long lo = T[i].PIndex, hi = T[i].PIndex + T[i].Count;
	long long begin = lo;
	for(int j = 0; j < 8 && begin < hi; j++)
	  {
	    // particles of daughter j: keys below the next daughter's first key (the last daughter takes the rest)
	    const long long end = j < 7 ? std::lower_bound(key + begin, key + hi, T[first + j + 1].StartKey) - key : hi;
	    gl_topnode &d = T[first + j];
	    if(j > 0) d.PIndex = (int) begin;	// domain.c:1563-1570: set when the scan steps into the daughter
	    double c = 0;
	    for(long long p = begin; p < end; p++) c += cost[p];
	    d.Cost = c;
	    d.Count = end - begin;
	    begin = end;
	  }
	for(int j = 0; j < 8; j++)
	  if(refine(first + j)) return 1;
      }
    return 0;
  }

  void number_leaves(int no, int &nleaves)	// domain_walktoptree, domain.c:1495
  {
    if(T[no].Daughter == -1)
      T[no].Leaf = nleaves++;
    else
      for(int j = 0; j < 8; j++) number_leaves(T[no].Daughter + j, nleaves);
  }
};
}				// namespace

extern "C" int gl_domain_top_tree(const gl_peanokey *Key, const double *costfactor, long long NumPart, int NTask, int multipledomains,
				  double topnodefactor, gl_topnode *nodes, int max_nodes, int *ntopnodes, int *ntopleaves)
{
  if(NumPart <= 0 || NumPart > 0x7fffffffLL || max_nodes < 1) return 2;
  std::vector<int> order((size_t) NumPart);
  sort_block(Key, 0, NumPart, order.data());	// stable, like mysort_domain (domain.c:1758)
  std::vector<unsigned long long> ks((size_t) NumPart);
  std::vector<double> cs((size_t) NumPart);
  HostPool::get().run(NumPart, 65536, [&](long long, long long a, long long b) {
    for(long long p = a; p < b; p++)
      {
	ks[(size_t) p] = Key[order[(size_t) p]];
	cs[(size_t) p] = costfactor[order[(size_t) p]];
      }
  });
  double gravcost = 0;		// domain.c:407-414, in P[] order
  for(long long i = 0; i < NumPart; i++) gravcost += costfactor[i];
  TopBuild B;
  B.T = nodes;
  B.n = 1;
  B.max_nodes = max_nodes;
  B.key = ks.data();
  B.cost = cs.data();
  B.costlimit = gravcost / (topnodefactor * multipledomains * NTask);	// domain.c:1772-1773 (one rank: totals = local sums)
  B.countlimit = (double) NumPart / (topnodefactor * multipledomains * NTask);
  nodes[0].Daughter = -1;
  nodes[0].Parent = -1;
  nodes[0].Size = ((gl_peanokey) 1) << 63;	// PEANOCELLS, 3 * BITS_PER_DIMENSION
  nodes[0].StartKey = 0;
  nodes[0].PIndex = 0;
  nodes[0].Count = NumPart;
  nodes[0].Cost = gravcost;
  nodes[0].Leaf = 0;
  if(B.refine(0)) return 1;	// "We are out of Topnodes" - the caller repeats with more room (domain.c:246-269)
  // domain.c:1871-1900: leaves that are still too heavy get daughters with an eighth of their count and cost
  for(int i = 0; i < B.n; i++)
    if(nodes[i].Daughter < 0 && (nodes[i].Count > B.countlimit || nodes[i].Cost > B.costlimit) && nodes[i].Size > 1)
      {
	if(B.n + 8 > max_nodes) return 1;
	nodes[i].Daughter = B.n;
	for(int j = 0; j < 8; j++)
	  {
	    gl_topnode &d = nodes[B.n + j];
	    d.Size = nodes[i].Size >> 3;
	    d.Count = nodes[i].Count / 8;
	    d.Cost = nodes[i].Cost / 8;
	    d.Daughter = -1;
	    d.Parent = i;
	    d.StartKey = nodes[i].StartKey + j * d.Size;
	    d.PIndex = 0;
	    d.Leaf = 0;
	  }
	B.n += 8;
      }
  int nleaves = 0;
  B.number_leaves(0, nleaves);
  if(ntopnodes) *ntopnodes = B.n;
  if(ntopleaves) *ntopleaves = nleaves;
  return 0;
}

// domain_findExtent (domain.c:1972-2014) + the key loop of domain_determineTopTree (domain.c:1748-1755) on all host
// cores: minimum / maximum are exact in any order, the cube is formed with the reference's expressions, and every key is
// peano_hilbert_key((int) ((Pos - DomainCorner) * DomainFac), ...) (peano.c:300).
extern "C" int gl_gadget_domain_keys(const gl_gadget_layout *lay, const void *Pv, long long NumPart, double DomainCorner[3],
				     double DomainCenter[3], double *DomainLen, double *DomainFac, gl_peanokey *Key)
{
  if(NumPart <= 0) return 1;
  const gl_gadget_layout L = *lay;
  const char *P = (const char *) Pv;
  const long long STRETCH = 65536, nst = (NumPart + STRETCH - 1) / STRETCH;
  std::vector<double> lo((size_t) (3 * nst), 1e300), hi((size_t) (3 * nst), -1e300);
  HostPool &pool = HostPool::get();
  pool.run(NumPart, STRETCH, [&](long long k, long long a, long long b) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for(long long i = a; i < b; i++)
      {
	const double *x = (const double *) (P + i * L.p_stride + L.p_Pos);
	for(int j = 0; j < 3; j++)
	  {
	    if(mn[j] > x[j]) mn[j] = x[j];
	    if(mx[j] < x[j]) mx[j] = x[j];
	  }
      }
    for(int j = 0; j < 3; j++)
      {
	lo[(size_t) (3 * k + j)] = mn[j];
	hi[(size_t) (3 * k + j)] = mx[j];
      }
  });
  double xmin[3] = {1e300, 1e300, 1e300}, xmax[3] = {-1e300, -1e300, -1e300};
  for(long long k = 0; k < nst; k++)
    for(int j = 0; j < 3; j++)
      {
	if(xmin[j] > lo[(size_t) (3 * k + j)]) xmin[j] = lo[(size_t) (3 * k + j)];
	if(xmax[j] < hi[(size_t) (3 * k + j)]) xmax[j] = hi[(size_t) (3 * k + j)];
      }
  double len = 0;
  for(int j = 0; j < 3; j++)
    if(xmax[j] - xmin[j] > len) len = xmax[j] - xmin[j];
  len *= 1.001;
  for(int j = 0; j < 3; j++)
    {
      DomainCenter[j] = 0.5 * (xmin[j] + xmax[j]);
      DomainCorner[j] = 0.5 * (xmin[j] + xmax[j]) - 0.5 * len;
    }
  *DomainLen = len;
  const double fac = 1.0 / len * (((gl_peanokey) 1) << 21);	// BITS_PER_DIMENSION
  *DomainFac = fac;
  const double c0 = DomainCorner[0], c1 = DomainCorner[1], c2 = DomainCorner[2];
  pool.run(NumPart, 8192, [=](long long, long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	const double *x = (const double *) (P + i * L.p_stride + L.p_Pos);
	Key[i] = gl_peano_hilbert_key((int) ((x[0] - c0) * fac), (int) ((x[1] - c1) * fac), (int) ((x[2] - c2) * fac), 21);
      }
  });
  return 0;
}

// domain_particle_costfactor (domain.c:378-384) for every particle
extern "C" int gl_gadget_costfactors(const gl_gadget_layout *lay, const void *Pv, long long NumPart, double *cost)
{
  const gl_gadget_layout L = *lay;
  const char *P = (const char *) Pv;
  HostPool::get().run(NumPart, 16384, [=](long long, long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	const char *pi = P + i * L.p_stride;
	const int tb = *(const short *) (pi + L.p_TimeBin);
	const double g = 1.0 + *(const float *) (pi + L.p_GravCost);
	cost[i] = tb ? g / (1 << tb) : g / (1 << 29);	// TIMEBASE
      }
  });
  return 0;
}

// ---- the multi-rank half of domain_decompose() (domain.c:470-487) ------------------------------------------------------
// These run identically on every rank of the reference, on leaf tables domain_sumCost has summed over the ranks
// (domain.c:1950-1960) - so they are plain functions of those tables and of NTask, restated with the reference's
// statements in the reference's order.  qsort() is the C library's, as in the reference: equal loads fall the same way.
namespace
{
struct LoadOrigin
{
  double load;
  int origin;
};
struct Segment
{
  int task, start, end;
};
int cmp_loadorigin(const void *a, const void *b)
{
  const double x = ((const LoadOrigin *) a)->load, y = ((const LoadOrigin *) b)->load;
  return x < y ? -1 : (x > y ? +1 : 0);
}
int cmp_segments(const void *a, const void *b)
{
  const int x = ((const Segment *) a)->task, y = ((const Segment *) b)->task;
  return x < y ? -1 : (x > y ? +1 : 0);
}
// domain_findSplit_work_balanced (domain.c:1075) / domain_findSplit_load_balanced (:1225): ncpu stretches of leaves of
// about equal sum
template <class T> void find_split(const T *v, int ncpu, int ndomain, int *start_list, int *end_list)
{
  double sum = 0;
  for(int i = 0; i < ndomain; i++) sum += v[i];
  const double avg = sum / ncpu;
  double before = 0, avg_before = 0;
  int start = 0;
  for(int i = 0; i < ncpu; i++)
    {
      double w = 0;
      int end = start;
      w += v[end];
      while((w + before < avg + avg_before) || (i == ncpu - 1 && end < ndomain - 1))
	{
	  if((ndomain - end) > (ncpu - i))
	    end++;
	  else
	    break;
	  w += v[end];
	}
      start_list[i] = start;
      end_list[i] = end;
      before += w;
      avg_before += avg;
      start = end + 1;
    }
}
// domain_assign_load_or_work_balanced (domain.c:1153): the MULTIPLEDOMAINS * NTask stretches are paired, heaviest with
// lightest, until NTask groups are left
void assign_balanced(int mode, int NTask, int md, const float *work, const int *count, int *start_list, int *end_list, int *domain_task)
{
  const int nseg = md * NTask;
  std::vector<Segment> seg((size_t) nseg);
  std::vector<LoadOrigin> dom((size_t) nseg);
  std::vector<int> target((size_t) nseg);
  for(int n = 0; n < nseg; n++) seg[(size_t) n].task = n;
  int ndomains = nseg;
  while(ndomains > NTask)
    {
      for(int i = 0; i < ndomains; i++)
	{
	  dom[(size_t) i].load = 0;
	  dom[(size_t) i].origin = i;
	}
      for(int n = 0; n < nseg; n++)
	for(int i = start_list[n]; i <= end_list[n]; i++)
	  {
	    if(mode == 1)
	      dom[(size_t) seg[(size_t) n].task].load += count[i];
	    else
	      dom[(size_t) seg[(size_t) n].task].load += work[i];
	  }
      qsort(dom.data(), (size_t) ndomains, sizeof(LoadOrigin), cmp_loadorigin);
      for(int i = 0; i < ndomains / 2; i++)
	{
	  target[(size_t) dom[(size_t) i].origin] = i;
	  target[(size_t) dom[(size_t) (ndomains - 1 - i)].origin] = i;
	}
      for(int n = 0; n < nseg; n++) seg[(size_t) n].task = target[(size_t) seg[(size_t) n].task];
      ndomains /= 2;
    }
  for(int n = 0; n < nseg; n++)
    {
      seg[(size_t) n].start = start_list[n];
      seg[(size_t) n].end = end_list[n];
    }
  qsort(seg.data(), (size_t) nseg, sizeof(Segment), cmp_segments);
  for(int n = 0; n < nseg; n++)
    {
      start_list[n] = seg[(size_t) n].start;
      end_list[n] = seg[(size_t) n].end;
      for(int i = start_list[n]; i <= end_list[n]; i++) domain_task[i] = seg[(size_t) n].task;
    }
}
// domain_check_memory_bound (domain.c:595)
int memory_bound(int NTask, int md, const float *work, const int *count, const int *countsph, const int *start_list, const int *end_list,
		 long long max_load, long long max_load_sph, int *list_load, int *list_loadsph, double *list_work)
{
  int max_l = 0, max_s = 0;
  for(int ta = 0; ta < NTask; ta++)
    {
      int load = 0, sphload = 0;
      double w = 0;
      for(int m = 0; m < md; m++)
	for(int i = start_list[ta * md + m]; i <= end_list[ta * md + m]; i++)
	  {
	    load += count[i];
	    sphload += countsph ? countsph[i] : 0;
	    w += work[i];
	  }
      if(list_load) list_load[ta] = load;
      if(list_loadsph) list_loadsph[ta] = sphload;
      if(list_work) list_work[ta] = w;
      if(load > max_l) max_l = load;
      if(sphload > max_s) max_s = sphload;
    }
  return (max_l > max_load || max_s > max_load_sph) ? 1 : 0;
}
}	// namespace

extern "C" int gl_domain_assign(int NTask, int multipledomains, int ntopleaves, const float *leaf_work, const int *leaf_count,
				const int *leaf_count_sph, long long max_load, long long max_load_sph, int *domain_task, int *start_list,
				int *end_list, int *list_load, int *list_loadsph, double *list_work)
{
  const int md = multipledomains, nseg = md * NTask;
  // the pairing halves the number of groups: it ends at NTask only when md is a power of two (the reference's
  // MULTIPLEDOMAINS = 16); fewer leaves than stretches is endrun(112) (domain.c:1912)
  if(NTask < 1 || md < 1 || (md & (md - 1)) || ntopleaves < nseg || !leaf_work || !leaf_count || !domain_task || !start_list || !end_list) return 3;
  find_split(leaf_work, nseg, ntopleaves, start_list, end_list);
  assign_balanced(1, NTask, md, leaf_work, leaf_count, start_list, end_list, domain_task);
  if(!memory_bound(NTask, md, leaf_work, leaf_count, leaf_count_sph, start_list, end_list, max_load, max_load_sph, list_load, list_loadsph, list_work))
    return 0;
  // "the optimum balanced solution violates memory constraint, let's try something different" (domain.c:476-480)
  find_split(leaf_count, nseg, ntopleaves, start_list, end_list);
  assign_balanced(0, NTask, md, leaf_work, leaf_count, start_list, end_list, domain_task);
  return memory_bound(NTask, md, leaf_work, leaf_count, leaf_count_sph, start_list, end_list, max_load, max_load_sph, list_load, list_loadsph, list_work) ? 2 : 1;
}

// domain_countToGo (domain.c:1274-1305, the part in front of the memory-limited iteration): where every local particle
// goes under DomainTask[], and how many leave for each rank
extern "C" int gl_domain_count_to_go(const gl_topnode *T, const gl_peanokey *Key, const int *type, long long NumPart, const int *domain_task,
				     int ThisTask, int NTask, int *to_go, int *to_go_sph, int *dest)
{
  if(!T || !Key || !domain_task || !to_go || NTask < 1 || ThisTask < 0 || ThisTask >= NTask) return 3;
  for(int n = 0; n < NTask; n++)
    {
      to_go[n] = 0;
      if(to_go_sph) to_go_sph[n] = 0;
    }
  for(long long n = 0; n < NumPart; n++)
    {
      int no = 0;
      while(T[no].Daughter >= 0) no = T[no].Daughter + (int) ((Key[n] - T[no].StartKey) / (T[no].Size >> 3));
      no = T[no].Leaf;
      const int ta = domain_task[no];
      if(ta < 0 || ta >= NTask) return 3;
      if(dest) dest[n] = ta;
      if(ta != ThisTask)
	{
	  to_go[ta] += 1;
	  if(to_go_sph && type && type[n] == 0) to_go_sph[ta] += 1;
	}
    }
  return 0;
}

// domain_sumCost (domain.c:1920-1966: work and count of every top-level leaf; the work is a 4-byte float accumulated in
// P[] order) followed by domain_findSplit_work_balanced (domain.c:1075-1113, all speed factors 1): ncpu consecutive
// stretches of leaves of about equal work.
extern "C" int gl_domain_split(const gl_topnode *T, int ntopleaves, const gl_peanokey *Key, const double *costfactor,
			       long long NumPart, int ncpu, int *start_list, int *end_list, float *leaf_work, int *leaf_count)
{
  if(ntopleaves < ncpu || ncpu < 1) return 1;	// endrun(112), domain.c:1912
  std::vector<float> work((size_t) ntopleaves, 0.0f);
  std::vector<int> count((size_t) ntopleaves, 0);
  for(long long n = 0; n < NumPart; n++)
    {
      int no = 0;
      while(T[no].Daughter >= 0) no = T[no].Daughter + (int) ((Key[n] - T[no].StartKey) / (T[no].Size >> 3));
      no = T[no].Leaf;
      work[(size_t) no] += costfactor[n];
      count[(size_t) no] += 1;
    }
  double w = 0;
  for(int i = 0; i < ntopleaves; i++) w += work[(size_t) i];
  const double workavg = w / ncpu;
  double work_before = 0, workavg_before = 0;
  int start = 0;
  for(int i = 0; i < ncpu; i++)
    {
      double wk = 0;
      int end = start;
      wk += work[(size_t) end] / 1.0;
      while((wk + work_before < workavg + workavg_before) || (i == ncpu - 1 && end < ntopleaves - 1))
	{
	  if((ntopleaves - end) > (ncpu - i))
	    end++;
	  else
	    break;
	  wk += work[(size_t) end] / 1.0;
	}
      start_list[i] = start;
      end_list[i] = end;
      work_before += wk;
      workavg_before += workavg;
      start = end + 1;
    }
  if(leaf_work) memcpy(leaf_work, work.data(), sizeof(float) * (size_t) ntopleaves);
  if(leaf_count) memcpy(leaf_count, count.data(), sizeof(int) * (size_t) ntopleaves);
  return 0;
}
