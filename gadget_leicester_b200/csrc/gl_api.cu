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
  cudaEventCreateWithFlags(&c->ev_blk, cudaEventBlockingSync | cudaEventDisableTiming);
  for(int i = 0; i < 4; i++) cudaEventCreateWithFlags(&c->ev_res[i], cudaEventBlockingSync | cudaEventDisableTiming);
  *out = c;
  return 0;
}

extern "C" void gl_destroy(gl_ctx *c)
{
  if(!c) return;
  cudaSetDevice(c->device);
  c->fire_post_launch();
  for(auto &t : c->push_threads)
    if(t.joinable()) t.join();
  cudaStreamSynchronize(c->stream);
  delete c->comm;
  c->comm = nullptr;
  gll_release(c);
  c->c_key.release(); c->c_level.release(); c->c_sm.release(); c->c_aux.release(); c->s_lev.release(); c->let_hsml_backup.release();
  c->in_velpred.release();
  c->run_err.release();
  c->run_hist.release();
  c->run_list.release();
  c->run_flag.release();
  c->hull_u64.release();
  c->hull_i32.release();
  c->run_id.release();
  if(c->run_pin) cudaFreeHost(c->run_pin);
  c->dma_rows.release();
  if(c->ev_blk) cudaEventDestroy(c->ev_blk);
  for(int i = 0; i < 4; i++)
    if(c->ev_res[i]) cudaEventDestroy(c->ev_res[i]);
  c->in_pos.release(); c->in_vel.release(); c->in_mass.release(); c->in_hsml.release(); c->in_oldacc.release();
  c->in_entropy.release(); c->in_dtentropy.release(); c->in_type.release(); c->in_timebin.release();
  c->in_tibegstep.release(); c->key_m.release(); c->key_m_sorted.release(); c->key_ph.release(); c->idx.release();
  c->perm.release(); c->inv.release(); c->ph_order.release(); c->cub_tmp.release(); c->dom_dev.release(); c->s_pm.release(); c->s_vh.release();
  c->s_type.release(); c->lev_c.release(); c->ncnt.release(); c->noff.release(); c->n_first.release(); c->n_end.release();
  c->n_parent.release(); c->n_level.release(); c->lev_list.release(); c->n_sm.release(); c->n_cl.release();
  c->n_aux.release(); c->p_father.release(); c->scan_tmp_i.release(); c->e_next.release(); c->e_A.release();
  c->e_F0.release(); c->e_F1.release(); c->e_F2.release(); c->grav_scale.release();
  c->e_B.release(); c->e_C.release(); c->e_hmax.release(); c->tgt_in.release(); c->tgt_sorted.release();
  c->tgt_sph.release(); c->tgt_flag.release(); c->tgt_gas.release(); c->o_acc.release(); c->o_oldacc.release();
  c->o_numngb.release(); c->o_density.release(); c->o_dhsml.release(); c->o_divvel.release(); c->o_curlvel.release();
  c->o_pressure.release(); c->o_sinkrho.release(); c->o_sinkentr.release(); c->o_sinkvel.release(); c->o_hacc.release();
  c->o_pot.release();
  c->o_dtentr.release(); c->o_maxsig.release(); c->o_gravcost.release(); c->o_denspairs.release();
  c->o_hydropairs.release(); c->d_counter.release(); c->h_q.release(); c->h_r.release();
  c->tgt_gas_part.release(); c->tgt_part_flag.release();
  c->sg_active.release(); c->sg_list.release(); c->sg_cnt.release(); c->sg_flag.release(); c->sg_scratch.release(); c->sg_vals.release();
  c->sg_claim.release(); c->sg_keys.release(); c->sg_off.release(); c->sg_pflag.release(); c->sg_slot.release(); c->sg_plist.release();
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

// a new particle set: own = all of it until a multi-rank build appends imports; a target list the multi-rank build had
// installed by itself goes with the old set
void gl_fresh_particle_set(gl_ctx *c)
{
  c->let_gap_hulls = false;	// a new set comes in the caller's order again
  c->n_own = c->n;
  c->ncell = 0;
  c->nit = c->n;
  if(c->let_auto_targets)
    {
      c->n_tgt_in = -1;
      c->tgt_ordered = false;
      c->let_auto_targets = false;
    }
  if(c->comm) gll_set_full(c, false);
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
  if(c->res_count >= 0 && c->res_first + c->res_count > n)
    {				// a result range set for a larger particle set does not survive the new one
      c->res_first = 0;
      c->res_count = -1;
    }
  c->n = n;
  gl_fresh_particle_set(c);
  c->tree_valid = false;
  c->tree_flag = true;
  c->tgt_dirty = true;
  c->have_velpred = false;
  c->run_ready = false;
  c->run_have_id = false;
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

// ---------------------------------------------------------------- multi-GPU: load from all-gather buffers ----
#define GL_MAX_RANKS 16
struct gl_segments
{
  long long start[GL_MAX_RANKS + 1];
  int world;
};

__device__ __forceinline__ int seg_of(const gl_segments &sg, long long g)
{
  int r = 0;
  while(r + 1 < sg.world && g >= sg.start[r + 1]) r++;
  return r;
}

// Block layout (per rank, m = padded segment length), gravity-essential rows first so a gravity-only exchange can
// send a prefix: doubles [pos 3m | mass m | hsml m | oldacc m || vel 3m | entropy m | dtentropy m], ints [type m ||
// timebin m | ti_begstep m].  full = 0: only the part before "||" is present (rows_f = 6, rows_i = 1).
__global__ void k_unpack_state(long long n, gl_segments sg, long long m, int full, const double *__restrict__ fb, const int *__restrict__ ib,
			       double *__restrict__ pos, double *__restrict__ vel, double *__restrict__ mass,
			       double *__restrict__ hsml, double *__restrict__ oldacc, double *__restrict__ entropy,
			       double *__restrict__ dtentropy, int *__restrict__ type, int *__restrict__ timebin,
			       int *__restrict__ tibeg)
{
  const long long g = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(g >= n) return;
  const int r = seg_of(sg, g);
  const long long j = g - sg.start[r];
  const double *f = fb + (long long) r * (full ? 11 : 6) * m;
  const int *q = ib + (long long) r * (full ? 3 : 1) * m;
  pos[3 * g] = f[3 * j]; pos[3 * g + 1] = f[3 * j + 1]; pos[3 * g + 2] = f[3 * j + 2];
  mass[g] = f[3 * m + j];
  hsml[g] = f[4 * m + j];
  oldacc[g] = f[5 * m + j];
  type[g] = q[j];
  if(full)
    {
      vel[3 * g] = f[6 * m + 3 * j]; vel[3 * g + 1] = f[6 * m + 3 * j + 1]; vel[3 * g + 2] = f[6 * m + 3 * j + 2];
      entropy[g] = f[9 * m + j];
      dtentropy[g] = f[10 * m + j];
      timebin[g] = q[m + j];
      tibeg[g] = q[2 * m + j];
    }
}

static int make_segments(gl_ctx *c, gl_segments &sg, int world, const long long *seg_start, long long n, long long m)
{
  if(world < 1 || world > GL_MAX_RANKS) return c->fail("gathered load: world size out of range (1..16)");
  sg.world = world;
  for(int r = 0; r <= world; r++) sg.start[r] = seg_start[r];
  if(sg.start[0] != 0 || sg.start[world] != n) return c->fail("gathered load: segments do not cover [0,n)");
  for(int r = 0; r < world; r++)
    if(sg.start[r + 1] < sg.start[r] || sg.start[r + 1] - sg.start[r] > m) return c->fail("gathered load: bad segment");
  return 0;
}

extern "C" int gl_load_particles_gathered(gl_ctx *c, long long n, int world, long long m, const long long *seg_start,
					   const double *f64buf, const int *i32buf, int full)
{
  cudaSetDevice(c->device);
  if(n <= 0) return c->fail("gl_load_particles_gathered: n <= 0");
  if(n >= (1ll << 27) + (1ll << 24)) return c->fail("gl_load_particles_gathered: n exceeds the 2^27 design limit of the index encoding");
  gl_segments sg;
  if(make_segments(c, sg, world, seg_start, n, m)) return 1;
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
  if(!full)
    {				// gravity-only exchange: the SPH-side inputs are not sent; keep them defined
      GL_CUDA(c, cudaMemsetAsync(c->in_vel.p, 0, 3 * n * sizeof(double), c->stream));
      GL_CUDA(c, cudaMemsetAsync(c->in_entropy.p, 0, n * sizeof(double), c->stream));
      GL_CUDA(c, cudaMemsetAsync(c->in_dtentropy.p, 0, n * sizeof(double), c->stream));
      GL_CUDA(c, cudaMemsetAsync(c->in_timebin.p, 0, n * sizeof(int), c->stream));
      GL_CUDA(c, cudaMemsetAsync(c->in_tibegstep.p, 0, n * sizeof(int), c->stream));
    }
  k_unpack_state<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, sg, m, full, f64buf, i32buf, c->in_pos.p, c->in_vel.p, c->in_mass.p, c->in_hsml.p,
							  c->in_oldacc.p, c->in_entropy.p, c->in_dtentropy.p, c->in_type.p,
							  c->in_timebin.p, c->in_tibegstep.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  if(c->res_count >= 0 && c->res_first + c->res_count > n)
    {
      c->res_first = 0;
      c->res_count = -1;
    }
  c->n = n;
  gl_fresh_particle_set(c);
  c->tree_valid = false;
  c->tree_flag = true;
  c->tgt_dirty = true;
  c->have_velpred = false;
  c->run_ready = false;
  c->run_have_id = false;
  if(!c->domain_imposed) c->have_domain = false;
  return 0;
}

// the SPH-side inputs that a split exchange delivers after the tree has been built: rank blocks of 5*m doubles
// [vel 3m | entropy m | dtentropy m] and 2*m ints [timebin m | ti_begstep m]
__global__ void k_unpack_late(long long n, gl_segments sg, long long m, const double *__restrict__ fb, const int *__restrict__ ib,
			      const int *__restrict__ inv, double4 *__restrict__ s_vh, double *__restrict__ vel,
			      double *__restrict__ entropy, double *__restrict__ dtentropy, int *__restrict__ timebin,
			      int *__restrict__ tibeg)
{
  const long long g = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(g >= n) return;
  const int r = seg_of(sg, g);
  const long long j = g - sg.start[r];
  const double *f = fb + (long long) r * 5 * m;
  const int *q = ib + (long long) r * 2 * m;
  const double vx = f[3 * j], vy = f[3 * j + 1], vz = f[3 * j + 2];
  vel[3 * g] = vx; vel[3 * g + 1] = vy; vel[3 * g + 2] = vz;
  entropy[g] = f[3 * m + j];
  dtentropy[g] = f[4 * m + j];
  timebin[g] = q[j];
  tibeg[g] = q[m + j];
  if(inv)
    {
      double4 *d = &s_vh[inv[g]];
      d->x = vx; d->y = vy; d->z = vz;
    }
}

extern "C" int gl_set_sph_inputs_gathered(gl_ctx *c, int world, long long m, const long long *seg_start, const double *f64rest,
					   const int *i32rest)
{
  cudaSetDevice(c->device);
  const long long n = c->n;
  if(n <= 0) return c->fail("gl_set_sph_inputs_gathered: no particles");
  gl_segments sg;
  if(make_segments(c, sg, world, seg_start, n, m)) return 1;
  k_unpack_late<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, sg, m, f64rest, i32rest, c->tree_valid ? c->inv.p : nullptr, c->s_vh.p,
							 c->in_vel.p, c->in_entropy.p, c->in_dtentropy.p, c->in_timebin.p, c->in_tibegstep.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  c->tgt_dirty = true;		// density_isactive (density.c:1035) reads TimeBin
  return 0;
}

__global__ void k_unpack_sph(long long n, gl_segments sg, long long m, const double *__restrict__ fb, const int *__restrict__ inv,
			     double4 *__restrict__ s_vh, double *__restrict__ hsml, double *__restrict__ density,
			     double *__restrict__ pressure, double *__restrict__ dhsml, double *__restrict__ divvel,
			     double *__restrict__ curlvel)
{
  const long long g = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(g >= n) return;
  const int r = seg_of(sg, g);
  const long long j = g - sg.start[r];
  const double *f = fb + (long long) r * 6 * m;
  const double h = f[j];
  hsml[g] = h;
  if(inv) s_vh[inv[g]].w = h;
  density[g] = f[m + j];
  pressure[g] = f[2 * m + j];
  dhsml[g] = f[3 * m + j];
  divvel[g] = f[4 * m + j];
  curlvel[g] = f[5 * m + j];
}

extern "C" int gl_set_sph_state_gathered(gl_ctx *c, int world, long long m, const long long *seg_start, const double *buf)
{
  cudaSetDevice(c->device);
  const long long n = c->n;
  if(n <= 0) return c->fail("gl_set_sph_state_gathered: no particles");
  gl_segments sg;
  if(make_segments(c, sg, world, seg_start, n, m)) return 1;
  GL_CUDA(c, c->o_density.ensure((size_t) n));
  GL_CUDA(c, c->o_pressure.ensure((size_t) n));
  GL_CUDA(c, c->o_dhsml.ensure((size_t) n));
  GL_CUDA(c, c->o_divvel.ensure((size_t) n));
  GL_CUDA(c, c->o_curlvel.ensure((size_t) n));
  k_unpack_sph<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, sg, m, buf, c->tree_valid ? c->inv.p : nullptr, c->s_vh.p, c->in_hsml.p,
							c->o_density.p, c->o_pressure.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
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

extern "C" void gl_set_tree_reconstruct_flag(gl_ctx *c, int flag)
{
  c->tree_flag = flag != 0;
  // multi-rank mode: a new decomposition / tree starts again with the pruned export (a full export after a cut cell was
  // met only lasts until then)
  if(flag && c->comm) gll_set_full(c, false);
}

// ---------------------------------------------------------------- targets ----
extern "C" int gl_set_targets_ordered(gl_ctx *c, const int *index, long long n, int where)
{
  if(gl_set_targets(c, index, n, where)) return 1;
  c->tgt_ordered = (c->n_tgt_in >= 0);
  return 0;
}

extern "C" int gl_set_targets(gl_ctx *c, const int *index, long long n, int where)
{
  cudaSetDevice(c->device);
  c->tgt_dirty = true;
  c->tgt_ordered = false;
  c->let_auto_targets = false;
  c->tgt_host.clear();
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
  if(!where && n > 0 && c->have_lay) c->tgt_host.assign(index, index + n);	// only the drop-in layer scatters on the host
  c->n_tgt_in = n;
  return 0;
}

// An index outside [0, n) is reported through *bad instead of being dereferenced (the reference would simply never
// list such a particle; an out-of-bounds device access would leave a sticky error and kill the context).
__global__ void k_mark(long long nt, const int *__restrict__ tin, const int *__restrict__ inv, int *__restrict__ flag, long long n,
		       int *__restrict__ bad)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nt) return;
  const int t = tin[i];
  if(t < 0 || t >= n)
    {
      *bad = 1;
      return;
    }
  flag[inv[t]] = 1;
}

// flags: bit0 target, bit1 SPH target (density_isactive, density.c:1035-1057), bit2 gas target
// thread k handles the k-th candidate: the k-th particle along the Peano-Hilbert curve (cand = ph_order), or the
// k-th entry of an ordered target list (cand built by k_list_candidates), so the selected lists keep that order
__global__ void k_classify(long long n, const int *__restrict__ cand, const int *__restrict__ s_type,
			   const int *__restrict__ in_timebin, const int *__restrict__ perm, int all,
			   const int *__restrict__ mark, int *__restrict__ flag, int *__restrict__ f_sph,
			   int *__restrict__ f_gas)
{
  long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k >= n) return;
  const int i = cand[k];
  int f = all ? 1 : mark[i];
  int ty = s_type[i];
  flag[k] = f;
  f_sph[k] = f && (ty == 0 || ty == 5 || ty == 2) && in_timebin[perm[i]] >= 0;
  f_gas[k] = f && ty == 0;
}

__global__ void k_list_candidates(long long nt, const int *__restrict__ tin, const int *__restrict__ inv, int *__restrict__ cand,
				  long long n, int *__restrict__ bad)
{
  long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k >= nt) return;
  const int t = tin[k];
  if(t < 0 || t >= n)
    {
      *bad = 1;
      cand[k] = 0;
      return;
    }
  cand[k] = inv[t];
}

static int select_flagged(gl_ctx *c, const int *items, const int *flags, int *out, long long n, long long *count)
{
  size_t tmp = 0;
  int *dcount = c->scan_tmp_i.p + 60;
  cub::DeviceSelect::Flagged(nullptr, tmp, items, flags, out, dcount, (int) n, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, items, flags, out, dcount, (int) n, c->stream));
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
  const int *cand;
  long long ncand = n;
  int *bad = c->scan_tmp_i.p + 62;
  const bool have_list = !all;
  if(have_list) GL_CUDA(c, cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
  if(!all && c->tgt_ordered)
    {
      // the caller's list order is kept: candidates = the list itself, every one of them a target
      ncand = c->n_tgt_in;
      if(ncand > n) return c->fail("targets: ordered list longer than the particle set");
      if(ncand > 0)
	{
	  k_list_candidates<<<gl_grid(ncand, 256), 256, 0, c->stream>>>(ncand, c->tgt_in.p, c->inv.p, mark, n, bad);
	  GL_LAUNCH_COUNT(c);
	}
      cand = mark;
      all = 1;
    }
  else
    {
      if(glk_ph_sequence(c)) return 1;	// the tree may have been built while an ordered list was set
      cand = c->ph_order.p;
      if(!all)
	{
	  GL_CUDA(c, cudaMemsetAsync(mark, 0, n * sizeof(int), c->stream));
	  if(c->n_tgt_in > 0)
	    {
	      k_mark<<<gl_grid(c->n_tgt_in, 256), 256, 0, c->stream>>>(c->n_tgt_in, c->tgt_in.p, c->inv.p, mark, n, bad);
	      GL_LAUNCH_COUNT(c);
	    }
	}
    }
  c->n_tgt = c->n_tgt_sph = c->n_tgt_gas = 0;
  if(have_list)
    {
      int hbad = 0;
      GL_CUDA(c, cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      if(hbad) return c->fail("targets: the list holds an index outside [0, n) of the loaded particle set");
    }
  if(ncand > 0)
    {
      k_classify<<<gl_grid(ncand, 256), 256, 0, c->stream>>>(ncand, cand, c->s_type.p, c->in_timebin.p, c->perm.p, all, mark, f0, f1, f2);
      GL_LAUNCH_COUNT(c);
      if(select_flagged(c, cand, f0, c->tgt_sorted.p, ncand, &c->n_tgt)) return 1;
      if(select_flagged(c, cand, f1, c->tgt_sph.p, ncand, &c->n_tgt_sph)) return 1;
      if(select_flagged(c, cand, f2, c->tgt_gas.p, ncand, &c->n_tgt_gas)) return 1;
    }
  c->tgt_dirty = false;
  return 0;
}

// Drop-in pipeline: reorder the gas target list into [targets with P[] index < split | the others] so hydro can run
// as two launches whose results are contiguous index ranges of the output arrays.  Both parts keep their
// Peano-Hilbert sequence (the second one reversed - DevicePartition writes rejected items back to front), so a
// warp's 32 targets stay a compact clump.
__global__ void k_part_flag(long long nt, const int *__restrict__ tgt, const int *__restrict__ perm, int split, int *__restrict__ flag)
{
  long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k < nt) flag[k] = perm[tgt[k]] < split;
}

int gl_split_gas_targets(gl_ctx *c, long long split, long long *n_lo)
{
  const long long nt = c->n_tgt_gas;
  *n_lo = 0;
  if(nt <= 0) return 0;
  GL_CUDA(c, c->tgt_gas_part.ensure((size_t) nt + 1));
  GL_CUDA(c, c->tgt_part_flag.ensure((size_t) nt + 1));
  GL_CUDA(c, c->scan_tmp_i.ensure(64));
  k_part_flag<<<gl_grid(nt, 256), 256, 0, c->stream>>>(nt, c->tgt_gas.p, c->perm.p, (int) split, c->tgt_part_flag.p);
  GL_LAUNCH_COUNT(c);
  size_t tmp = 0;
  int *dcount = c->scan_tmp_i.p + 61;
  cub::DevicePartition::Flagged(nullptr, tmp, c->tgt_gas.p, c->tgt_part_flag.p, c->tgt_gas_part.p, dcount, (int) nt, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DevicePartition::Flagged(c->cub_tmp.p, tmp, c->tgt_gas.p, c->tgt_part_flag.p, c->tgt_gas_part.p, dcount, (int) nt, c->stream));
  c->kernel_launches += 2;
  int h = 0;
  GL_CUDA(c, cudaMemcpyAsync(&h, dcount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_lo = h;
  return 0;
}

// ---------------------------------------------------------------- force entry points ----
// Multi-rank mode: a phase that met a cut cell (GL_ERR_LET_VIOLATION, see gl_let.cu) is repeated by ALL ranks with a
// full export - the decision is collective (one 8-byte all-gather per phase), because the rebuild is.
static int run_phase(gl_ctx *c, int (*fn)(gl_ctx *), const char *name, bool gravity, bool restore_hsml = false)
{
  if(!(c->comm && c->comm->world > 1)) return fn(c);
  for(int attempt = 0; attempt < 2; attempt++)
    {
      const int rc = fn(c);
      const bool viol = rc != 0 && c->err.find("GL_ERR_LET_VIOLATION") != std::string::npos;
      if(rc != 0 && !viol) return rc;
      int any = 0;
      if(gll_agree_violation(c, viol ? 1 : 0, &any)) return 1;
      if(!any) return 0;
      if(attempt == 1) break;
      if(getenv("GL_DEBUG_LET")) fprintf(stderr, "[gl] rank %d: %s met a cut cell somewhere - repeating with a full export\n", c->comm->rank, name);
      c->let_violations++;
      gll_set_full(c, true);
      c->tree_flag = true;
      c->tree_valid = false;
      if(!gravity)
	{
	  c->let_need_sph_exchange = true;
	  // density only: the repeated Hsml iteration starts from the values the first attempt started from
	  if(restore_hsml && c->let_hsml_backup.cap >= (size_t) c->n_own && c->n_own > 0)
	    GL_CUDA(c, cudaMemcpyAsync(c->in_hsml.p, c->let_hsml_backup.p, c->n_own * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	  if(glt_build(c)) return 1;
	}
    }
  return c->fail(std::string(name) + ": a cut cell was met although every particle was exported");
}

static int density_phase(gl_ctx *c)
{
  if(c->comm && c->comm->world > 1)
    {
      // a repeated density() must start its Hsml iteration from the same values as the first attempt
      if(!c->tree_valid || c->tree_flag)
	if(glt_build(c)) return 1;
      GL_CUDA(c, c->let_hsml_backup.ensure((size_t) c->n_own + 1));
      GL_CUDA(c, cudaMemcpyAsync(c->let_hsml_backup.p, c->in_hsml.p, c->n_own * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
  return gls_density(c);
}

extern "C" int gl_gravity_tree(gl_ctx *c)
{
  cudaSetDevice(c->device);
  if(c->comm && c->comm->world > 1)
    {
      c->let_want_grav = true;
      // the merged tree was selected for one opening criterion (and maybe without the gravity tests at all)
      if(c->tree_valid && !gll_built_for(c, c->par.ErrTolTheta != 0, true)) c->tree_flag = true;
    }
  bool rebuilt = c->tree_flag || !c->tree_valid;
  if(rebuilt) c->tgt_dirty = true;
  return run_phase(c, glg_gravity, "gravity_tree", true);
}

extern "C" int gl_compute_potential(gl_ctx *c)
{
  cudaSetDevice(c->device);
  if(c->comm && c->comm->world > 1) return c->fail("compute_potential: single rank only (all particles are targets, potential.c:54)");
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
  return run_phase(c, density_phase, "density", false, true);
}

extern "C" int gl_force_update_hmax(gl_ctx *c)
{
  cudaSetDevice(c->device);
  return glt_update_hmax(c);
}

static int hydro_phase(gl_ctx *c)
{
  // after a repeat the tree is new: the SphP state of the imports and the cells' hmax have to come again
  if(c->comm && c->comm->world > 1 && c->let_need_sph_exchange)
    if(glt_update_hmax(c)) return 1;
  return gls_hydro(c);
}

extern "C" int gl_hydro_force(gl_ctx *c)
{
  cudaSetDevice(c->device);
  return run_phase(c, hydro_phase, "hydro_force", false);
}

// ---- multi-rank mode ----
extern "C" int gl_comm_unique_id(void *id128)
{
  std::string err;
  if(gl_comm_nccl_unique_id(id128, err))
    {
      fprintf(stderr, "libgl_b200: %s\n", err.c_str());
      return 1;
    }
  return 0;
}

extern "C" int gl_comm_init(gl_ctx *c, int rank, int world, const void *id128)
{
  cudaSetDevice(c->device);
  if(c->comm) return c->fail("gl_comm_init: a communicator is attached already");
  if(world < 1 || world > GL_LET_MAXRANKS || rank < 0 || rank >= world) return c->fail("gl_comm_init: bad rank / world");
  std::string err;
  c->comm = gl_comm_make_nccl(rank, world, id128, c->device, err);
  if(!c->comm) return c->fail(err);
  return 0;
}

extern "C" void *gl_comm_local_group_create(int world) { return gl_comm_local_group_create_impl(world); }
extern "C" void gl_comm_local_group_destroy(void *group) { gl_comm_local_group_destroy_impl(group); }

extern "C" int gl_comm_init_local(gl_ctx *c, int rank, void *group)
{
  if(c->comm) return c->fail("gl_comm_init_local: a communicator is attached already");
  std::string err;
  c->comm = gl_comm_make_local(group, rank, err);
  if(!c->comm) return c->fail(err);
  return 0;
}

extern "C" int gl_comm_finalize(gl_ctx *c)
{
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  delete c->comm;
  c->comm = nullptr;
  return 0;
}

extern "C" int gl_comm_set_options(gl_ctx *c, int want_gravity, double search_padding)
{
  if(!c->comm) return c->fail("gl_comm_set_options: no communicator attached");
  if(search_padding < 1.0) return c->fail("gl_comm_set_options: search_padding < 1");
  c->let_want_grav = want_gravity != 0;
  gll_set_padding(c, search_padding);
  return 0;
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
  if(c->comm && c->comm->world > 1)
    {
      // the SPH phases are collective: a rank whose active list holds no SPH particle in this step (individual time
      // steps) still has to serve the others' neighbour searches
      int any = 0;
      if(gll_agree_violation(c, have_sph > 0 ? 1 : 0, &any)) return 1;
      have_sph = any;
    }
  if(have_sph > 0)
    {
      if(gl_density(c)) return 1;	// accel.c:84
      if(gl_force_update_hmax(c)) return 1;	// accel.c:96
      if(gl_hydro_force(c)) return 1;	// accel.c:106
    }
  return 0;
}

// ---------------------------------------------------------------- results ----
// cnt = elements for the whole particle set; with a result range set (gl_set_result_range) only the elements of
// particles [first, first+count) are copied, to the same positions of dst
template <class T> static int get(gl_ctx *c, T *dst, const DBuf<T> &src, size_t cnt, int where, const char *what)
{
  if(!dst) return 0;
  size_t off = 0;
  if(c->res_count >= 0 && c->n > 0)
    {
      const size_t w = cnt / (size_t) c->n;	// 1 or 3 values per particle
      off = (size_t) c->res_first * w;
      cnt = (size_t) c->res_count * w;
    }
  // (multi-rank mode: a later phase that had to be repeated with more imports makes c->n larger than it was when an
  // earlier phase sized its result arrays - only the requested range has to exist)
  if(src.cap < off + cnt) return c->fail(std::string("result not available: ") + what);
  if(cnt == 0) return 0;
  GL_CUDA(c, cudaMemcpyAsync(dst + off, src.p + off, cnt * sizeof(T), where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
			     c->stream));
  return 0;
}

extern "C" int gl_set_result_range(gl_ctx *c, long long first, long long count)
{
  if(count < 0)
    {
      c->res_first = 0;
      c->res_count = -1;
      return 0;
    }
  if(first < 0 || first + count > c->n) return c->fail("gl_set_result_range: range outside the particle set");
  c->res_first = first;
  c->res_count = count;
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

