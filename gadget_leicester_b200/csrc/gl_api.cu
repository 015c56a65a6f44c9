// gl_api.cu - extern "C" boundary of libgl_b200.so (see include/gl_b200.h for the contract).
#include "gl_internal.h"
#include <cub/cub.cuh>
#include <thread>
#include <chrono>
#include <algorithm>

gl_dparams gl_make_dparams(const gl_ctx *c);
template <class T> static int get(gl_ctx *c, T *dst, const DBuf<T> &src, size_t cnt, int where, const char *what);
static int finish_get(gl_ctx *c, int evslot);

extern "C" int gl_create(gl_ctx **out, int device, long long max_part)
{
  *out = nullptr;
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    {
      fprintf(stderr, "libgl_b200: no CUDA device visible - this library has no CPU fallback\n");
      return 1;
    }
  if(device < 0 || device >= ndev) return 1;
  if(cudaSetDevice(device) != cudaSuccess) return 1;
  gl_ctx *c = new gl_ctx();
  c->device = device;
  c->max_part = max_part;
  if(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
      delete c;
      return 1;
    }
  for(int i = 0; i < 16; i++) cudaEventCreate(&c->ev[i]);
  *out = c;
  return 0;
}

extern "C" void gl_destroy(gl_ctx *c)
{
  if(!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->in_pos.release(); c->in_vel.release(); c->in_mass.release(); c->in_hsml.release(); c->in_oldacc.release();
  c->in_entropy.release(); c->in_dtentropy.release(); c->in_type.release(); c->in_timebin.release();
  c->in_tibegstep.release(); c->key_m.release(); c->key_m_sorted.release(); c->key_ph.release(); c->idx.release();
  c->perm.release(); c->inv.release(); c->ph_order.release(); c->cub_tmp.release(); c->dom_dev.release(); c->s_pm.release(); c->s_vh.release();
  c->s_type.release(); c->lev_c.release(); c->ncnt.release(); c->noff.release(); c->n_first.release(); c->n_end.release();
  c->n_parent.release(); c->n_level.release(); c->lev_list.release(); c->n_sm.release(); c->n_cl.release();
  c->n_aux.release(); c->p_father.release(); c->scan_tmp_i.release(); c->e_next.release(); c->e_A.release();
  c->e_B.release(); c->e_C.release(); c->e_hmax.release(); c->tgt_in.release(); c->tgt_sorted.release();
  c->tgt_sph.release(); c->tgt_flag.release(); c->tgt_gas.release(); c->o_acc.release(); c->o_oldacc.release();
  c->o_numngb.release(); c->o_density.release(); c->o_dhsml.release(); c->o_divvel.release(); c->o_curlvel.release();
  c->o_pressure.release(); c->o_sinkrho.release(); c->o_sinkentr.release(); c->o_sinkvel.release(); c->o_hacc.release();
  c->o_pot.release();
  c->o_dtentr.release(); c->o_maxsig.release(); c->o_gravcost.release(); c->o_denspairs.release();
  c->o_hydropairs.release(); c->d_counter.release(); c->h_q.release(); c->h_r.release();
  c->e_rng.release(); c->n_child.release(); c->fat_p.release(); c->fat_i.release(); c->fat_state.release();
  c->fat_rec.release(); c->fat_tgt.release();
  if(c->pin) cudaFreeHost(c->pin);
  if(c->pin2) cudaFreeHost(c->pin2);
  for(int k = 0; k < 2; k++)
    if(c->copy_stream[k]) cudaStreamDestroy(c->copy_stream[k]);
  for(int i = 0; i < 16; i++) cudaEventDestroy(c->ev[i]);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char *gl_last_error(const gl_ctx *c) { return c ? c->err.c_str() : "null context"; }
extern "C" void *gl_stream(gl_ctx *c) { return (void *) c->stream; }

extern "C" int gl_set_params(gl_ctx *c, const gl_params *p)
{
  if(p->ComovingIntegrationOn) return c->fail("gl_set_params: ComovingIntegrationOn != 0 is not implemented");
  c->par = *p;
  c->have_par = true;
  return 0;
}

template <class T> static int put(gl_ctx *c, DBuf<T> &dst, const T *src, size_t cnt, int where, T fill)
{
  GL_CUDA(c, dst.ensure(cnt));
  if(src)
    GL_CUDA(c, cudaMemcpyAsync(dst.p, src, cnt * sizeof(T), where ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
			       c->stream));
  else if(fill == T(0))
    GL_CUDA(c, cudaMemsetAsync(dst.p, 0, cnt * sizeof(T), c->stream));
  return 0;
}

extern "C" int gl_load_particles(gl_ctx *c, long long n, const double *pos, const double *vel, const double *mass,
				 const int *type, const double *hsml, const double *oldacc, const int *timebin,
				 const int *ti_begstep, const double *entropy, const double *dtentropy, int where)
{
  cudaSetDevice(c->device);
  if(n <= 0) return c->fail("gl_load_particles: n <= 0");
  if(n >= (1ll << 27) + (1ll << 24)) return c->fail("gl_load_particles: n exceeds the 2^27 design limit of the index encoding");
  if(!pos || !vel || !mass || !type) return c->fail("gl_load_particles: pos/vel/mass/type are required");
  cudaEventRecord(c->ev[8], c->stream);
  if(put(c, c->in_pos, pos, 3 * (size_t) n, where, 0.0)) return 1;
  if(put(c, c->in_vel, vel, 3 * (size_t) n, where, 0.0)) return 1;
  if(put(c, c->in_mass, mass, (size_t) n, where, 0.0)) return 1;
  if(put(c, c->in_type, type, (size_t) n, where, 0)) return 1;
  if(put(c, c->in_hsml, hsml, (size_t) n, where, 0.0)) return 1;
  if(put(c, c->in_oldacc, oldacc, (size_t) n, where, 0.0)) return 1;
  if(put(c, c->in_timebin, timebin, (size_t) n, where, 0)) return 1;
  if(put(c, c->in_tibegstep, ti_begstep, (size_t) n, where, 0)) return 1;
  if(put(c, c->in_entropy, entropy, (size_t) n, where, 0.0)) return 1;
  if(put(c, c->in_dtentropy, dtentropy, (size_t) n, where, 0.0)) return 1;
  cudaEventRecord(c->ev[9], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]);
  c->st.ms_h2d = ms;
  c->n = n;
  c->tree_valid = false;
  c->tree_flag = true;
  c->tgt_dirty = true;
  if(!c->domain_imposed) c->have_domain = false;
  return 0;
}

extern "C" int gl_set_oldacc(gl_ctx *c, const double *oldacc, int where)
{
  if(c->n <= 0) return c->fail("gl_set_oldacc: no particles");
  return put(c, c->in_oldacc, oldacc, (size_t) c->n, where, 0.0);
}

__global__ void k_scatter_hsml(long long n, const int *__restrict__ inv, const double *__restrict__ hs, double4 *s_vh)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) s_vh[inv[i]].w = hs[i];
}

extern "C" int gl_set_hsml(gl_ctx *c, const double *hsml, int where)
{
  if(c->n <= 0) return c->fail("gl_set_hsml: no particles");
  if(put(c, c->in_hsml, hsml, (size_t) c->n, where, 0.0)) return 1;
  if(c->tree_valid)
    {
      k_scatter_hsml<<<gl_grid(c->n, 256), 256, 0, c->stream>>>(c->n, c->inv.p, c->in_hsml.p, c->s_vh.p);
      GL_LAUNCH_COUNT(c);
    }
  return 0;
}

extern "C" int gl_set_sph_state(gl_ctx *c, const double *density, const double *pressure, const double *dhsml,
				const double *divvel, const double *curlvel, int where)
{
  cudaSetDevice(c->device);
  if(c->n <= 0) return c->fail("gl_set_sph_state: no particles");
  const size_t n = (size_t) c->n;
  if(density && put(c, c->o_density, density, n, where, 0.0)) return 1;
  if(pressure && put(c, c->o_pressure, pressure, n, where, 0.0)) return 1;
  if(dhsml && put(c, c->o_dhsml, dhsml, n, where, 0.0)) return 1;
  if(divvel && put(c, c->o_divvel, divvel, n, where, 0.0)) return 1;
  if(curlvel && put(c, c->o_curlvel, curlvel, n, where, 0.0)) return 1;
  return 0;
}

extern "C" int gl_domain_find_extent(gl_ctx *c, double *out8)
{
  cudaSetDevice(c->device);
  if(c->n <= 0) return c->fail("gl_domain_find_extent: no particles");
  c->have_domain = false;
  if(glk_find_extent(c)) return 1;
  if(out8)
    {
      out8[0] = c->dom.len; out8[1] = c->dom.fac;
      for(int k = 0; k < 3; k++)
	{
	  out8[2 + k] = c->dom.corner[k];
	  out8[5 + k] = c->dom.center[k];
	}
    }
  return 0;
}

extern "C" int gl_peano_keys(gl_ctx *c, gl_peanokey *ph, gl_peanokey *mo, int where)
{
  cudaSetDevice(c->device);
  if(c->n <= 0) return c->fail("gl_peano_keys: no particles");
  return glk_export_keys(c, ph, mo, where);
}

extern "C" int gl_peano_hilbert_order(gl_ctx *c, long long n_gas, int *order, int where)
{
  cudaSetDevice(c->device);
  if(c->n <= 0) return c->fail("gl_peano_hilbert_order: no particles");
  return glk_ph_order(c, n_gas, order, where);
}

extern "C" int gl_force_treebuild(gl_ctx *c, long long *numnodes)
{
  cudaSetDevice(c->device);
  if(glt_build(c)) return 1;
  c->tgt_dirty = true;
  if(numnodes) *numnodes = c->nnodes;
  return 0;
}

extern "C" int gl_get_nodes(gl_ctx *c, double *table, long long cap)
{
  cudaSetDevice(c->device);
  return glt_get_nodes(c, table, cap);
}

extern "C" void gl_set_tree_reconstruct_flag(gl_ctx *c, int flag) { c->tree_flag = flag != 0; }

// ---------------------------------------------------------------- targets ----
extern "C" int gl_set_targets(gl_ctx *c, const int *index, long long n, int where)
{
  cudaSetDevice(c->device);
  c->tgt_dirty = true;
  if(n < 0 || !index)
    {
      c->n_tgt_in = -1;
      return 0;
    }
  GL_CUDA(c, c->tgt_in.ensure((size_t) n + 1));
  if(n > 0)
    GL_CUDA(c, cudaMemcpyAsync(c->tgt_in.p, index, n * sizeof(int), where ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
			       c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_tgt_in = n;
  return 0;
}

__global__ void k_mark(long long nt, const int *__restrict__ tin, const int *__restrict__ inv, int *__restrict__ flag)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < nt) flag[inv[tin[i]]] = 1;
}

// flags: bit0 target, bit1 SPH target (density_isactive, density.c:1035-1057), bit2 gas target
// thread k handles the k-th particle along the Peano-Hilbert curve, so the selected lists come out in PH order
__global__ void k_classify(long long n, const int *__restrict__ ph_order, const int *__restrict__ s_type,
			   const int *__restrict__ in_timebin, const int *__restrict__ perm, int all,
			   const int *__restrict__ mark, int *__restrict__ flag, int *__restrict__ f_sph,
			   int *__restrict__ f_gas)
{
  long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k >= n) return;
  const int i = ph_order[k];
  int f = all ? 1 : mark[i];
  int ty = s_type[i];
  flag[k] = f;
  f_sph[k] = f && (ty == 0 || ty == 5 || ty == 2) && in_timebin[perm[i]] >= 0;
  f_gas[k] = f && ty == 0;
}

static int select_flagged(gl_ctx *c, const int *flags, int *out, long long n, long long *count)
{
  const int *it = c->ph_order.p;
  size_t tmp = 0;
  int *dcount = c->scan_tmp_i.p + 60;
  cub::DeviceSelect::Flagged(nullptr, tmp, it, flags, out, dcount, (int) n, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, it, flags, out, dcount, (int) n, c->stream));
  c->kernel_launches += 2;
  int h = 0;
  GL_CUDA(c, cudaMemcpyAsync(&h, dcount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  *count = h;
  return 0;
}

int gl_prepare_targets(gl_ctx *c)
{
  if(!c->tgt_dirty) return 0;
  if(!c->tree_valid) return c->fail("targets: no tree");
  long long n = c->n;
  GL_CUDA(c, c->tgt_flag.ensure(4 * (size_t) n + 4));
  GL_CUDA(c, c->tgt_sorted.ensure((size_t) n + 1));
  GL_CUDA(c, c->tgt_sph.ensure((size_t) n + 1));
  GL_CUDA(c, c->tgt_gas.ensure((size_t) n + 1));
  GL_CUDA(c, c->scan_tmp_i.ensure(64));
  int *f0 = c->tgt_flag.p, *f1 = f0 + n, *f2 = f1 + n, *mark = f2 + n;
  int all = c->n_tgt_in < 0;
  if(!all)
    {
      GL_CUDA(c, cudaMemsetAsync(mark, 0, n * sizeof(int), c->stream));
      if(c->n_tgt_in > 0)
	{
	  k_mark<<<gl_grid(c->n_tgt_in, 256), 256, 0, c->stream>>>(c->n_tgt_in, c->tgt_in.p, c->inv.p, mark);
	  GL_LAUNCH_COUNT(c);
	}
    }
  k_classify<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->ph_order.p, c->s_type.p, c->in_timebin.p, c->perm.p, all, mark,
						      f0, f1, f2);
  GL_LAUNCH_COUNT(c);
  if(select_flagged(c, f0, c->tgt_sorted.p, n, &c->n_tgt)) return 1;
  if(select_flagged(c, f1, c->tgt_sph.p, n, &c->n_tgt_sph)) return 1;
  if(select_flagged(c, f2, c->tgt_gas.p, n, &c->n_tgt_gas)) return 1;
  c->tgt_dirty = false;
  return 0;
}

// ---------------------------------------------------------------- force entry points ----
extern "C" int gl_gravity_tree(gl_ctx *c)
{
  cudaSetDevice(c->device);
  bool rebuilt = c->tree_flag || !c->tree_valid;
  if(rebuilt) c->tgt_dirty = true;
  return glg_gravity(c);
}

extern "C" int gl_compute_potential(gl_ctx *c)
{
  cudaSetDevice(c->device);
  return glg_potential(c);
}

extern "C" int gl_get_potential(gl_ctx *c, double *potential, int where)
{
  cudaSetDevice(c->device);
  cudaEventRecord(c->ev[10], c->stream);
  if(get(c, potential, c->o_pot, (size_t) c->n, where, "Potential")) return 1;
  return finish_get(c, 10);
}

extern "C" int gl_density(gl_ctx *c)
{
  cudaSetDevice(c->device);
  return gls_density(c);
}

extern "C" int gl_force_update_hmax(gl_ctx *c)
{
  cudaSetDevice(c->device);
  return glt_update_hmax(c);
}

extern "C" int gl_hydro_force(gl_ctx *c)
{
  cudaSetDevice(c->device);
  return gls_hydro(c);
}

extern "C" int gl_compute_accelerations(gl_ctx *c, int mode)
{
  // accel.c:27-106 (force part).  mode is passed through for signature parity; the reference uses it
  // only for diagnostics output.
  (void) mode;
  cudaSetDevice(c->device);
  const bool first = (c->par.TypeOfOpeningCriterion == 1 && c->par.Ti_Current == 0 && c->par.ErrTolTheta != 0);
  if(gl_gravity_tree(c)) return 1;	// accel.c:61
  if(first)
    if(gl_gravity_tree(c)) return 1;	// accel.c:63-67: second pass so the relative criterion has an OldAcc
  long long have_sph = 0;
  if(gl_prepare_targets(c)) return 1;
  have_sph = c->n_tgt_sph;
  if(have_sph > 0)
    {
      if(gl_density(c)) return 1;	// accel.c:84
      if(gl_force_update_hmax(c)) return 1;	// accel.c:96
      if(gl_hydro_force(c)) return 1;	// accel.c:106
    }
  return 0;
}

// ---------------------------------------------------------------- results ----
template <class T> static int get(gl_ctx *c, T *dst, const DBuf<T> &src, size_t cnt, int where, const char *what)
{
  if(!dst) return 0;
  if(src.cap < cnt) return c->fail(std::string("result not available: ") + what);
  GL_CUDA(c, cudaMemcpyAsync(dst, src.p, cnt * sizeof(T), where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
			     c->stream));
  return 0;
}

static int finish_get(gl_ctx *c, int evslot)
{
  cudaEventRecord(c->ev[evslot + 1], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[evslot], c->ev[evslot + 1]);
  c->st.ms_d2h += ms;
  return 0;
}

extern "C" int gl_get_gravity(gl_ctx *c, double *acc3, double *oldacc, float *gravcost, int where)
{
  cudaSetDevice(c->device);
  size_t n = (size_t) c->n;
  cudaEventRecord(c->ev[10], c->stream);
  if(get(c, acc3, c->o_acc, 3 * n, where, "GravAccel")) return 1;
  if(get(c, oldacc, c->o_oldacc, n, where, "OldAcc")) return 1;
  if(get(c, gravcost, c->o_gravcost, n, where, "GravCost")) return 1;
  return finish_get(c, 10);
}

extern "C" int gl_get_density(gl_ctx *c, double *hsml, double *numngb, double *density, double *dhsml, double *divvel,
			      double *curlvel, double *pressure, double *sink_density, double *sink_entropy,
			      double *sink_gasvel3, int where)
{
  cudaSetDevice(c->device);
  size_t n = (size_t) c->n;
  cudaEventRecord(c->ev[10], c->stream);
  if(get(c, hsml, c->in_hsml, n, where, "Hsml")) return 1;
  if(get(c, numngb, c->o_numngb, n, where, "NumNgb")) return 1;
  if(get(c, density, c->o_density, n, where, "Density")) return 1;
  if(get(c, dhsml, c->o_dhsml, n, where, "DhsmlDensityFactor")) return 1;
  if(get(c, divvel, c->o_divvel, n, where, "DivVel")) return 1;
  if(get(c, curlvel, c->o_curlvel, n, where, "CurlVel")) return 1;
  if(get(c, pressure, c->o_pressure, n, where, "Pressure")) return 1;
  if(get(c, sink_density, c->o_sinkrho, n, where, "sink density")) return 1;
  if(get(c, sink_entropy, c->o_sinkentr, n, where, "sink entropy")) return 1;
  if(get(c, sink_gasvel3, c->o_sinkvel, 3 * n, where, "sink gas velocity")) return 1;
  return finish_get(c, 10);
}

extern "C" int gl_get_hydro(gl_ctx *c, double *hacc3, double *dtentropy, double *maxsig, int where)
{
  cudaSetDevice(c->device);
  size_t n = (size_t) c->n;
  cudaEventRecord(c->ev[10], c->stream);
  if(get(c, hacc3, c->o_hacc, 3 * n, where, "HydroAccel")) return 1;
  if(get(c, dtentropy, c->o_dtentr, n, where, "DtEntropy")) return 1;
  if(get(c, maxsig, c->o_maxsig, n, where, "MaxSignalVel")) return 1;
  return finish_get(c, 10);
}

extern "C" int gl_get_pair_counts(gl_ctx *c, int *dens_pairs, int *hydro_pairs, int where)
{
  cudaSetDevice(c->device);
  size_t n = (size_t) c->n;
  cudaEventRecord(c->ev[10], c->stream);
  if(get(c, dens_pairs, c->o_denspairs, n, where, "density pair counts")) return 1;
  if(get(c, hydro_pairs, c->o_hydropairs, n, where, "hydro pair counts")) return 1;
  return finish_get(c, 10);
}

extern "C" int gl_get_stats(gl_ctx *c, gl_stats *out)
{
  c->st.kernel_launches = c->kernel_launches;
  *out = c->st;
  return 0;
}

extern "C" void gl_reset_launch_count(gl_ctx *c)
{
  c->kernel_launches = 0;
  c->st.ms_d2h = 0;
}

// ---------------------------------------------------------------- gadget AoS layer ----
extern "C" int gl_gadget_bind(gl_ctx *c, const gl_gadget_layout *lay)
{
  if(lay->p_stride <= 0 || lay->sph_stride <= 0) return c->fail("gl_gadget_bind: bad strides");
  c->lay = *lay;
  c->have_lay = true;
  return 0;
}

static int ensure_pinned(gl_ctx *c, size_t bytes)
{
  if(bytes <= c->pin_cap) return 0;
  if(c->pin) cudaFreeHost(c->pin);
  if(c->pin2) cudaFreeHost(c->pin2);
  for(int k = 0; k < 2; k++)
    if(c->copy_stream[k]) cudaStreamDestroy(c->copy_stream[k]);
  c->pin = nullptr;
  c->pin_cap = 0;
  GL_CUDA(c, cudaMallocHost(&c->pin, bytes + bytes / 8));
  c->pin_cap = bytes + bytes / 8;
  return 0;
}

template <class F> static void parallel_for(long long n, F f)
{
  unsigned hw = std::thread::hardware_concurrency();
  int nth = (int) std::min<long long>(std::max(1u, std::min(hw, 32u)), (n + 65535) / 65536);
  if(nth <= 1)
    {
      f(0, n);
      return;
    }
  std::vector<std::thread> th;
  long long chunk = (n + nth - 1) / nth;
  for(int t = 0; t < nth; t++)
    {
      long long a = t * chunk, b = std::min(n, a + chunk);
      if(a >= b) break;
      th.emplace_back([=]() { f(a, b); });
    }
  for(auto &x : th) x.join();
}

// Staging layout (all in one pinned block, SoA): pos 3n | vel 3n | mass n | hsml n | oldacc n | entropy n |
// dtentropy n | type n(int) | timebin n(int) | tibegstep n(int)
extern "C" int gl_gadget_pull(gl_ctx *c, const void *Pv, const void *SphPv, long long NumPart, long long N_gas)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_pull: gl_gadget_bind not called");
  const long long n = NumPart;
  if(n <= 0) return c->fail("gl_gadget_pull: NumPart <= 0");
  const size_t bytes = (size_t) n * (11 * sizeof(double) + 3 * sizeof(int));
  if(ensure_pinned(c, bytes)) return 1;
  double *pos = (double *) c->pin, *vel = pos + 3 * n, *mass = vel + 3 * n, *hsml = mass + n, *oldacc = hsml + n,
    *entropy = oldacc + n, *dtentropy = entropy + n;
  int *type = (int *) (dtentropy + n), *timebin = type + n, *tibeg = timebin + n;
  const gl_gadget_layout L = c->lay;
  const char *P = (const char *) Pv, *S = (const char *) SphPv;
  parallel_for(n, [=](long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	const char *pi = P + i * L.p_stride;
	const double *x = (const double *) (pi + L.p_Pos);
	pos[3 * i] = x[0]; pos[3 * i + 1] = x[1]; pos[3 * i + 2] = x[2];
	mass[i] = *(const double *) (pi + L.p_Mass);
	hsml[i] = *(const double *) (pi + L.p_Hsml);
	oldacc[i] = *(const double *) (pi + L.p_OldAcc);
	const int ty = *(const short *) (pi + L.p_Type);
	type[i] = ty;
	timebin[i] = *(const short *) (pi + L.p_TimeBin);
	tibeg[i] = *(const int *) (pi + L.p_Ti_begstep);
	if(ty == 0 && i < N_gas)
	  {
	    const char *si = S + i * L.sph_stride;
	    const double *v = (const double *) (si + L.s_VelPred);
	    vel[3 * i] = v[0]; vel[3 * i + 1] = v[1]; vel[3 * i + 2] = v[2];
	    entropy[i] = *(const double *) (si + L.s_Entropy);
	    dtentropy[i] = *(const double *) (si + L.s_DtEntropy);
	  }
	else
	  {
	    const double *v = (const double *) (pi + L.p_Vel);
	    vel[3 * i] = v[0]; vel[3 * i + 1] = v[1]; vel[3 * i + 2] = v[2];
	    entropy[i] = 0;
	    dtentropy[i] = 0;
	  }
      }
  });
  c->g_numpart = NumPart;
  c->g_ngas = N_gas;
  return gl_load_particles(c, n, pos, vel, mass, type, hsml, oldacc, timebin, tibeg, entropy, dtentropy, 0);
}

// ---- result scatter: D2H of compact SoA results on `st`, then a multi-threaded scatter into P[] / SphP[] ----
static int push_gravity_on(gl_ctx *c, cudaStream_t st, double *stage, char *P)
{
  cudaSetDevice(c->device);
  const long long n = c->n;
  const gl_gadget_layout L = c->lay;
  double *acc = stage, *old = acc + 3 * n;
  float *cost = (float *) (old + n);
  GL_CUDA(c, cudaMemcpyAsync(acc, c->o_acc.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaMemcpyAsync(old, c->o_oldacc.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaMemcpyAsync(cost, c->o_gravcost.p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaStreamSynchronize(st));
  parallel_for(n, [=](long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	char *pi = P + i * L.p_stride;
	double *g = (double *) (pi + L.p_GravAccel);
	g[0] = acc[3 * i]; g[1] = acc[3 * i + 1]; g[2] = acc[3 * i + 2];
	*(double *) (pi + L.p_OldAcc) = old[i];
	*(float *) (pi + L.p_GravCost) = cost[i];
      }
  });
  return 0;
}

static int push_density_on(gl_ctx *c, cudaStream_t st, double *stage, char *P, char *S)
{
  cudaSetDevice(c->device);
  const long long n = c->n, N_gas = c->g_ngas;
  const gl_gadget_layout L = c->lay;
  double *hs = stage, *nn = hs + n, *de = nn + n, *dh = de + n, *dv = dh + n, *cu = dv + n, *pr = cu + n, *sr = pr + n,
    *se = sr + n, *sv = se + n;
  const double *src[9] = { c->in_hsml.p, c->o_numngb.p, c->o_density.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p,
    c->o_pressure.p, c->o_sinkrho.p, c->o_sinkentr.p };
  double *dst[9] = { hs, nn, de, dh, dv, cu, pr, sr, se };
  const bool have_other = (n > N_gas);	// sink / dust fields only exist for non-gas targets
  for(int k = 0; k < (have_other ? 9 : 7); k++)
    GL_CUDA(c, cudaMemcpyAsync(dst[k], src[k], n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if(have_other) GL_CUDA(c, cudaMemcpyAsync(sv, c->o_sinkvel.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaStreamSynchronize(st));
  parallel_for(n, [=](long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	char *pi = P + i * L.p_stride;
	const int ty = *(const short *) (pi + L.p_Type);
	if(!(ty == 0 || ty == 5 || ty == 2)) continue;
	*(double *) (pi + L.p_Hsml) = hs[i];
	*(double *) (pi + L.p_NumNgb) = nn[i];
	*(unsigned int *) (pi + L.p_SwallowID) = 0;	// density.c:130
	if(ty == 0 && i < N_gas)
	  {
	    char *si = S + i * L.sph_stride;
	    *(double *) (si + L.s_Density) = de[i];
	    *(double *) (si + L.s_DhsmlDensityFactor) = dh[i];
	    *(double *) (si + L.s_DivVel) = dv[i];
	    *(double *) (si + L.s_CurlVel) = cu[i];
	    *(double *) (si + L.s_Pressure) = pr[i];
	  }
	else if(ty == 5)
	  {
	    *(double *) (pi + L.p_SinkDensity) = sr[i];
	    *(double *) (pi + L.p_SinkEntropy) = se[i];
	    double *g = (double *) (pi + L.p_SinkGasVel);
	    g[0] = sv[3 * i]; g[1] = sv[3 * i + 1]; g[2] = sv[3 * i + 2];
	  }
	else if(ty == 2)
	  {
	    *(double *) (pi + L.p_DustDensity) = sr[i];
	    *(double *) (pi + L.p_DustEntropy) = se[i];
	    double *g = (double *) (pi + L.p_DustGasVel);
	    g[0] = sv[3 * i]; g[1] = sv[3 * i + 1]; g[2] = sv[3 * i + 2];
	  }
      }
  });
  return 0;
}

static int push_hydro_on(gl_ctx *c, cudaStream_t st, double *stage, char *P, char *S)
{
  cudaSetDevice(c->device);
  const long long n = c->n, N_gas = c->g_ngas;
  const gl_gadget_layout L = c->lay;
  double *ha = stage, *dt = ha + 3 * n, *ms = dt + n;
  GL_CUDA(c, cudaMemcpyAsync(ha, c->o_hacc.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaMemcpyAsync(dt, c->o_dtentr.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaMemcpyAsync(ms, c->o_maxsig.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GL_CUDA(c, cudaStreamSynchronize(st));
  parallel_for(std::min(n, N_gas), [=](long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	const char *pi = P + i * L.p_stride;
	if(*(const short *) (pi + L.p_Type) != 0) continue;
	char *si = S + i * L.sph_stride;
	double *g = (double *) (si + L.s_HydroAccel);
	g[0] = ha[3 * i]; g[1] = ha[3 * i + 1]; g[2] = ha[3 * i + 2];
	*(double *) (si + L.s_DtEntropy) = dt[i];
	*(double *) (si + L.s_MaxSignalVel) = ms[i];
      }
  });
  return 0;
}

static int ensure_pinned2(gl_ctx *c, size_t bytes)
{
  if(bytes <= c->pin2_cap) return 0;
  if(c->pin2) cudaFreeHost(c->pin2);
  c->pin2 = nullptr;
  c->pin2_cap = 0;
  GL_CUDA(c, cudaMallocHost(&c->pin2, bytes + bytes / 8));
  c->pin2_cap = bytes + bytes / 8;
  return 0;
}

extern "C" int gl_gadget_push(gl_ctx *c, void *Pv, void *SphPv, int what)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_push: gl_gadget_bind not called");
  const long long n = c->n;
  if(ensure_pinned2(c, (size_t) n * 24 * sizeof(double))) return 1;
  double *stage = (double *) c->pin2;
  if((what & 1) && push_gravity_on(c, c->stream, stage, (char *) Pv)) return 1;
  if((what & 2) && push_density_on(c, c->stream, stage, (char *) Pv, (char *) SphPv)) return 1;
  if((what & 4) && push_hydro_on(c, c->stream, stage, (char *) Pv, (char *) SphPv)) return 1;
  return 0;
}

// ---- fused scatters of the pipelined drop-in call: each struct array is touched exactly once ----
// P[]: gravity results + the P-side density results, in one pass (struct lines 1, 2 and 8 of 9)
static void scatter_P_fused(gl_ctx *c, char *P, const double *stA, const double *stB, bool sph)
{
  const long long n = c->n, N_gas = c->g_ngas;
  const gl_gadget_layout L = c->lay;
  const double *acc = stA, *old = acc + 3 * n;
  const float *cost = (const float *) (old + n);
  const double *hs = stB, *nn = hs + n, *sr = nn + 6 * n, *se = sr + n, *sv = se + n;
  const int *type = (const int *) ((const double *) c->pin + 11 * n);	// Type column of the pull staging (still valid)
  const bool have_other = (n > N_gas);
  parallel_for(n, [=](long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	char *pi = P + i * L.p_stride;
	double *g = (double *) (pi + L.p_GravAccel);
	g[0] = acc[3 * i]; g[1] = acc[3 * i + 1]; g[2] = acc[3 * i + 2];
	*(double *) (pi + L.p_OldAcc) = old[i];
	*(float *) (pi + L.p_GravCost) = cost[i];
	if(!sph) continue;
	const int ty = type[i];
	if(!(ty == 0 || ty == 5 || ty == 2)) continue;
	*(double *) (pi + L.p_Hsml) = hs[i];
	*(double *) (pi + L.p_NumNgb) = nn[i];
	*(unsigned int *) (pi + L.p_SwallowID) = 0;	// density.c:130
	if(ty == 5 && have_other)
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

// SphP[]: density + hydro results in one pass (struct lines 0, 1, 2 of 5)
static void scatter_SphP_fused(gl_ctx *c, char *S, const double *stB, const double *stC)
{
  const long long n = c->n, N_gas = c->g_ngas;
  const gl_gadget_layout L = c->lay;
  const double *de = stB + 2 * n, *dh = de + n, *dv = dh + n, *cu = dv + n, *pr = cu + n;
  const double *ha = stC, *dt = ha + 3 * n, *ms = dt + n;
  const int *type = (const int *) ((const double *) c->pin + 11 * n);
  parallel_for(std::min(n, N_gas), [=](long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	if(type[i] != 0) continue;
	char *si = S + i * L.sph_stride;
	*(double *) (si + L.s_Pressure) = pr[i];
	*(double *) (si + L.s_MaxSignalVel) = ms[i];
	*(double *) (si + L.s_Density) = de[i];
	*(double *) (si + L.s_DtEntropy) = dt[i];
	double *g = (double *) (si + L.s_HydroAccel);
	g[0] = ha[3 * i]; g[1] = ha[3 * i + 1]; g[2] = ha[3 * i + 2];
	*(double *) (si + L.s_DhsmlDensityFactor) = dh[i];
	*(double *) (si + L.s_DivVel) = dv[i];
	*(double *) (si + L.s_CurlVel) = cu[i];
      }
  });
}

// The drop-in compute_accelerations (accel.c:27) on the reference's host arrays, software-pipelined:
//   GPU:   gravity | density            | hmax + hydro        |
//   PCIe:          | D2H gravity results| D2H density results | D2H hydro results
//   host:                               | scatter into P[]    | scatter into SphP[]
// Every struct array is read once (pull) and written once; the P[] scatter hides behind the hydro kernel.
extern "C" int gl_gadget_compute_accelerations(gl_ctx *c, void *Pv, void *SphPv, long long NumPart, long long N_gas, int mode)
{
  (void) mode;
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_compute_accelerations: gl_gadget_bind not called");
  const bool trace = getenv("GL_DEBUG_E2E") != nullptr;
  auto now_ms = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now_ms();
  double t_pull = 0, t_grav = 0, t_dens = 0, t_hyd = 0, t_pP = 0;
  if(gl_gadget_pull(c, Pv, SphPv, NumPart, N_gas)) return 1;
  t_pull = now_ms();
  const long long n = c->n;
  if(ensure_pinned2(c, (size_t) n * 24 * sizeof(double))) return 1;
  if(!c->copy_stream[0])
    for(int k = 0; k < 2; k++) GL_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream[k], cudaStreamNonBlocking));
  double *stA = (double *) c->pin2, *stB = stA + 5 * n, *stC = stB + 13 * n;	// 5n | 13n | 5n doubles
  char *P = (char *) Pv, *S = (char *) SphPv;
  const bool first = (c->par.TypeOfOpeningCriterion == 1 && c->par.Ti_Current == 0 && c->par.ErrTolTheta != 0);
  if(gl_gravity_tree(c)) return 1;	// accel.c:61
  if(first)
    if(gl_gravity_tree(c)) return 1;	// accel.c:63-67
  t_grav = now_ms();
  // gravity results start crossing PCIe while density runs
  cudaStream_t cs0 = c->copy_stream[0], cs1 = c->copy_stream[1];
  GL_CUDA(c, cudaMemcpyAsync(stA, c->o_acc.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, cs0));
  GL_CUDA(c, cudaMemcpyAsync(stA + 3 * n, c->o_oldacc.p, n * sizeof(double), cudaMemcpyDeviceToHost, cs0));
  GL_CUDA(c, cudaMemcpyAsync(stA + 4 * n, c->o_gravcost.p, n * sizeof(float), cudaMemcpyDeviceToHost, cs0));
  if(gl_prepare_targets(c)) return 1;
  const bool sph = c->n_tgt_sph > 0;
  int rc = 0;
  std::thread tP;
  if(sph)
    {
      rc = gl_density(c);	// accel.c:84
      t_dens = now_ms();
      if(rc == 0)
	{
	  const double *src[9] = { c->in_hsml.p, c->o_numngb.p, c->o_density.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p,
	    c->o_pressure.p, c->o_sinkrho.p, c->o_sinkentr.p };
	  const bool have_other = (n > N_gas);
	  for(int k = 0; k < (have_other ? 9 : 7); k++)
	    GL_CUDA(c, cudaMemcpyAsync(stB + k * n, src[k], n * sizeof(double), cudaMemcpyDeviceToHost, cs1));
	  if(have_other) GL_CUDA(c, cudaMemcpyAsync(stB + 9 * n, c->o_sinkvel.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, cs1));
	  tP = std::thread([&]() {
	    cudaSetDevice(c->device);
	    cudaStreamSynchronize(cs0);
	    cudaStreamSynchronize(cs1);
	    scatter_P_fused(c, P, stA, stB, true);
	    t_pP = now_ms();
	  });
	  rc = gl_force_update_hmax(c);	// accel.c:96
	  if(rc == 0) rc = gl_hydro_force(c);	// accel.c:106
	  t_hyd = now_ms();
	  if(rc == 0)
	    {
	      GL_CUDA(c, cudaMemcpyAsync(stC, c->o_hacc.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	      GL_CUDA(c, cudaMemcpyAsync(stC + 3 * n, c->o_dtentr.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	      GL_CUDA(c, cudaMemcpyAsync(stC + 4 * n, c->o_maxsig.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	      GL_CUDA(c, cudaStreamSynchronize(c->stream));
	      GL_CUDA(c, cudaStreamSynchronize(cs1));
	      scatter_SphP_fused(c, S, stB, stC);
	    }
	}
    }
  else
    {
      GL_CUDA(c, cudaStreamSynchronize(cs0));
      scatter_P_fused(c, P, stA, stB, false);
    }
  if(tP.joinable()) tP.join();
  if(trace)
    fprintf(stderr, "[gl] e2e ms since entry: pull %.1f | gravity done %.1f | density done %.1f | hydro done %.1f | P[] scatter done %.1f | exit %.1f\n",
	    t_pull - t_start, t_grav - t_start, t_dens - t_start, t_hyd - t_start, t_pP - t_start, now_ms() - t_start);
  return rc;
}
