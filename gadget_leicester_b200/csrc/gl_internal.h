// gl_internal.h - shared declarations of libgl_b200 (not part of the public ABI)
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <functional>
#include <thread>
#include <atomic>
#include "../../include/gl_b200.h"
#include "gl_comm.h"

#define GL_BITS 21			// BITS_PER_DIMENSION, allvars.h:58
#define GL_MAXITER 150			// MAXITER, allvars.h:139
#define GL_WARPS_PER_BLOCK 4
#define GL_BLOCK (32 * GL_WARPS_PER_BLOCK)

struct gl_error_t { int code; };

#define GL_CUDA(ctx, call)                                                                       \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if(e__ != cudaSuccess) {                                                                     \
      (ctx)->fail(std::string(#call) + ": " + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                  std::to_string(__LINE__));                                                     \
      return 1;                                                                                  \
    }                                                                                            \
  } while(0)

// growable device buffer
template <class T> struct DBuf
{
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n)
  {
    if(n <= cap) return cudaSuccess;
    if(p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 16 + 64;
    cudaError_t e = cudaMalloc((void **) &p, want * sizeof(T));
    if(e == cudaSuccess) cap = want;
    return e;
  }
  // grow without losing the contents (multi-rank mode appends imported particles behind the caller's own, and a
  // rebuild with more imports must not lose results of an earlier phase).  Rare: waits for the device.
  cudaError_t ensure_keep(size_t n)
  {
    if(n <= cap) return cudaSuccess;
    size_t want = n + n / 8 + 64;
    T *q = nullptr;
    cudaError_t e = cudaMalloc((void **) &q, want * sizeof(T));
    if(e != cudaSuccess) return e;
    if(p)
      {
	cudaDeviceSynchronize();
	e = cudaMemcpy(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice);
	cudaFree(p);
	if(e != cudaSuccess)
	  {
	    cudaFree(q);
	    p = nullptr;
	    cap = 0;
	    return e;
	  }
      }
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release()
  {
    if(p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// domain cube, exactly the quantities of domain.c:2004-2013
struct gl_domain
{
  double len, fac;
  double corner[3], center[3];
};

// device-side parameter block (copied by value into kernels)
struct gl_dparams
{
  double theta, errtolforceacc;
  double soft[6];		// ForceSoftening
  double softtab[6];		// SofteningTable
  double G, lambda_fac;
  double desnumngb, maxdev, bhfac, mingashsml;
  double artvisc, timebase_interval;
  int ti_current;
};

struct gl_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  gl_params par{};
  bool have_par = false;
  long long max_part = 0;
  long long n = 0;		// loaded particles
  bool tree_flag = true;	// TreeReconstructFlag
  bool tree_valid = false;
  bool have_domain = false, domain_imposed = false;
  gl_domain dom{};
  long long kernel_launches = 0;
  gl_stats st{};
  // once-per-device set-up that must not be process-global: a process may hold contexts on several GPUs
  bool tables_up = false, attr_density = false, attr_hydro = false;

  // ---- multi-rank mode (gl_comm_init): the caller loads its OWN particles; the tree build imports, per other rank,
  // the locally-essential set that rank selects for us - particles and whole cells ("cut cells": moments only, never
  // opened).  Input arrays then hold own particles [0, n_own) followed by the imported ones [n_own, n); the imported
  // cells are items n .. nit-1 of the sorted order.  Single rank: n_own = n, nit = n.
  gl_comm *comm = nullptr;
  long long n_own = 0, ncell = 0, nit = 0;
  DBuf<unsigned long long> c_key;	// cut cells: lowest Morton key of the cell
  DBuf<signed char> c_level;	// its level (1..21)
  DBuf<double4> c_sm, c_aux;	// (s, mass), (maxsoft, hmax, vmax, -)
  DBuf<signed char> s_lev;	// tree order: level of a cut-cell item, 22 for a particle
  struct gl_let_state *let = nullptr;
  bool let_auto_targets = false;	// the target list was set by the multi-rank build (own particles), not by the caller
  long long let_violations = 0;
  bool let_need_sph_exchange = false;	// the tree was rebuilt after density(): the imports' SphP state has to come again
  bool let_want_grav = true;	// the export test includes the gravity opening criteria (gl_comm_set_options)
  DBuf<double> let_hsml_backup;	// own Hsml before density(): a repeated density() must start from the same values
  // ---- input-order particle state (index = P[] index) ----
  DBuf<double> in_pos, in_vel, in_mass, in_hsml, in_oldacc, in_entropy, in_dtentropy;
  // drop-in layer only: SphP[].VelPred of the gas block, kept apart from P[].Vel (node vmax uses P.Vel,
  // forcetree.c:684; the SPH sums use VelPred, density.c:772 / hydra.c:962).  Applied to s_vh after every build.
  DBuf<double> in_velpred;
  long long n_velpred = 0;
  bool have_velpred = false;
  DBuf<int> in_type, in_timebin, in_tibegstep;
  // ---- keys / sort ----
  DBuf<unsigned long long> key_m, key_m_sorted, key_ph;
  DBuf<int> idx, perm;		// perm[sorted] = input index
  DBuf<int> inv;		// inv[input] = sorted index
  DBuf<int> ph_order;		// ph_order[k] = tree (Morton) index of the k-th particle along the Peano-Hilbert curve
  DBuf<unsigned char> cub_tmp;
  DBuf<double> dom_dev;		// 8 doubles + scratch for the extent reduction
  // ---- sorted particle state (tree order = Morton order) ----
  DBuf<double4> s_pm;		// x,y,z,mass
  DBuf<double4> s_vh;		// vx,vy,vz,hsml
  DBuf<int> s_type;
  // ---- tree ----
  DBuf<signed char> lev_c;	// c[i] common level of sorted keys i-1,i ; c[0]=c[n]=-1
  DBuf<int> ncnt, noff;		// nodes starting at particle a, exclusive scan
  long long nnodes = 0, nentries = 0;
  DBuf<int> n_first, n_end, n_parent;
  DBuf<signed char> n_level;
  DBuf<int> lev_list;		// node ids grouped by level
  int lev_start[GL_BITS + 2]{};
  DBuf<double4> n_sm;		// s[3], mass
  DBuf<double4> n_cl;		// center[3], len
  DBuf<double4> n_aux;		// maxsoft, hmax, vmax, divVmax
  DBuf<int> n_child;		// 8 child codes per node (see k_moments)
  DBuf<int> p_father;		// Father[] analogue: node id containing sorted particle directly
  DBuf<int> scan_tmp_i;
  // ---- walk entries (DFS order, nodes and particle leaves interleaved) ----
  DBuf<int> e_next;		// node: skip entry; particle: -1-(p*8+type)
  DBuf<double4> e_A;		// node: s,mass ; particle: pos,mass
  DBuf<double4> e_B;		// node: center,len ; particle: soft,0,0,-1
  DBuf<double2> e_C;		// node: maxsoft, mass*len*len ; particle: unused
  DBuf<double> e_hmax;		// node: hmax (for the symmetric search)
  DBuf<int2> e_rng;		// node: [first,end) range of tree-ordered particles below it
  // FP32 walk records (gl_gravity.cu): positions as an exact grid part + a small remainder, in units that make the
  // domain cube [1,2) wide and the total mass [1,2) (powers of two: the opening decisions do not change)
  DBuf<float4> e_F0;		// hi(x,y,z) of s / pos, mass
  DBuf<float4> e_F1;		// lo(x,y,z), node: len ; particle: softening length
  DBuf<float4> e_F2;		// node: centre - s (x,y,z), maxsoft
  DBuf<double> grav_scale;	// [0] length scale S, [1] mass scale Ms (device-written after the moments)
  // ---- targets ----
  DBuf<int> tgt_in;		// user-supplied P[] indices (or empty = all)
  std::vector<int> tgt_host;	// host copy of the same list (the drop-in scatter only touches targets)
  long long n_tgt_in = -1;
  long long res_first = 0, res_count = -1;	// gl_set_result_range: gl_get_* copy only this particle range (-1 = all)
  bool tgt_ordered = false;	// the list is already in a spatially coherent order (gl_set_targets_ordered): no PH sort
  bool have_ph_order = false;
  DBuf<int> tgt_sorted;		// sorted-order indices of the gravity targets
  DBuf<int> tgt_sph;		// sorted-order indices of SPH targets (gas/sink/dust)
  DBuf<int> tgt_flag;
  long long n_tgt = 0, n_tgt_sph = 0, n_tgt_gas = 0;
  bool tgt_dirty = true;
  DBuf<int> tgt_gas;
  // ---- outputs, input order ----
  DBuf<double> o_acc, o_oldacc, o_numngb, o_density, o_dhsml, o_divvel, o_curlvel, o_pressure, o_sinkrho,
    o_sinkentr, o_sinkvel, o_hacc, o_dtentr, o_maxsig;
  DBuf<float> o_gravcost;
  DBuf<double> o_pot;
  DBuf<int> o_denspairs, o_hydropairs;
  DBuf<unsigned long long> d_counter;	// [0] interactions, [1] max passes, [2] error flag
  // ---- hydro source records (sorted order) ----
  DBuf<double4> h_q;		// rho, P/rho^2, P/rho^2*Omega, soundspeed
  DBuf<double4> h_r;		// f (Balsara), timestep, 1/h, 1/h^4
  // ---- fat density targets (block-per-target path) ----
  DBuf<int> fat_p, fat_i;
  DBuf<double4> fat_state, fat_rec;
  DBuf<double> fat_tgt;
  // ---- sub-grid neighbour loops (gl_subgrid.cu) ----
  gl_subgrid_params subgrid{};
  bool have_subgrid = false;
  DBuf<int> sg_active, sg_list, sg_cnt, sg_pflag, sg_slot, sg_plist;
  DBuf<unsigned char> sg_flag;
  DBuf<double> sg_scratch, sg_vals;
  DBuf<unsigned long long> sg_claim, sg_keys;
  DBuf<long long> sg_off;
  long long sg_pairs_seen = 0;	// (grain, gas) pairs of the last gl_dust_drag: sizes the next call's list
  // ---- device-resident integration loop (gl_resident.cu) ----
  DBuf<int> run_err, run_hist, run_list, run_flag;
  DBuf<unsigned long long> hull_u64;	// multi-rank resident loop: order-independent key hulls (glk_ph_gap_hulls)
  DBuf<int> hull_i32;
  bool let_gap_hulls = false;
  DBuf<int> run_id;		// particle ids travelling with gl_domain_exchange
  bool run_have_id = false;
  int run_ti = 0;		// the time every particle is drifted to
  bool run_ready = false;
  void *run_pin = nullptr;	// page-locked staging of gl_gadget_run_steps
  size_t run_pin_cap = 0;
  // ---- gadget AoS layer ----
  gl_gadget_layout lay{};
  bool have_lay = false;
  long long g_numpart = 0, g_ngas = 0;	// particles of the host arrays bound to this context (one rank's P[] / SphP[])
  long long g_off = 0;		// their first global index: 0 on one GPU, the segment start when the sources were all-gathered
  const int *g_type = nullptr;	// host copy of P[].Type for them (a column of the pack staging)
  void *pin = nullptr, *pin2 = nullptr;
  size_t pin_cap = 0, pin2_cap = 0;
  cudaStream_t copy_stream[2] = { nullptr, nullptr };
  // strided copy-engine path of the drop-in layer (pinned / registered host arrays only)
  DBuf<char> dma_rows;		// H2D: the leading bytes of P[] structs, one 128-byte-pitched row per particle

  cudaEvent_t ev[16]{};
  // the drop-in call keeps every host core but one busy with pack/scatter work: its long waits must sleep, not spin
  // drop-in pipeline: called once, right after the next phase's main kernel has been launched.  Result copies are
  // enqueued from here because a kernel submitted BEHIND a batch of D2H copies was measured to start only when
  // the copies had finished (2.9 ms per step), although they run on another stream.
  std::function<void()> post_launch;
  std::vector<std::thread> push_threads;	// gl_gadget_push_segment_async scatter threads, joined by gl_gadget_push_wait
  std::atomic<int> push_failed{0};	// written by the scatter threads
  void fire_post_launch()
  {
    if(post_launch)
      {
	std::function<void()> f = std::move(post_launch);
	post_launch = nullptr;
	f();
      }
  }
  // hydro targets split by P[] index (drop-in pipeline: results of the first part cross PCIe while the second runs)
  DBuf<int> tgt_gas_part;
  DBuf<int> tgt_part_flag;
  cudaEvent_t ev_blk = nullptr;
  cudaEvent_t ev_res[4] = { nullptr, nullptr, nullptr, nullptr };	// "result block has landed in host memory" (blocking-sync events)
  bool blocking_sync = false;
  cudaError_t wait_stream()
  {
    if(!blocking_sync || !ev_blk) return cudaStreamSynchronize(stream);
    cudaError_t e = cudaEventRecord(ev_blk, stream);
    if(e != cudaSuccess) return e;
    return cudaEventSynchronize(ev_blk);
  }

  int fail(const std::string &m)
  {
    err = m;
    return 1;
  }
};

// gl_keys.cu
int glk_find_extent(gl_ctx *c);
int glk_keys_and_sort(gl_ctx *c);
int glk_export_keys(gl_ctx *c, gl_peanokey *ph, gl_peanokey *mo, int where);
int glk_ph_order(gl_ctx *c, long long n_gas, int *order, int where);
int glk_apply_velpred(gl_ctx *c);
int glk_ph_sequence(gl_ctx *c);
int glk_local_minmax(gl_ctx *c, double *dev6);	// min[3], max[3] of the loaded positions, left on the device
void glk_set_domain_minmax(gl_ctx *c, const double lo[3], const double hi[3]);
// Peano-Hilbert key hull {min, max} of each of nchunks stretches [bounds[b], bounds[b+1]) of the own particles (input order)
int glk_ph_chunk_hulls(gl_ctx *c, const long long *bounds_dev, int nchunks, unsigned long long *out);
// the same cover from the SORTED own keys, cut at the largest gaps: independent of the caller's order (resident loop)
int glk_ph_gap_hulls(gl_ctx *c, int nchunks, unsigned long long *out);
// per node of the current tree: 1 if no other rank has a particle inside the cell (its Peano-Hilbert key range
// misses every key hull of every other rank); rank r's nhull hulls start at ranges + r * stride (device)
int glk_node_complete(gl_ctx *c, const unsigned long long *ranges, int world, int nhull, long long stride, int me, unsigned char *complete);
// gl_tree.cu
int glt_build(gl_ctx *c);
int glt_update_hmax(gl_ctx *c);
int glt_get_nodes(gl_ctx *c, double *table, long long cap);
// gl_gravity.cu
int glg_gravity(gl_ctx *c);
int glg_potential(gl_ctx *c);
// gl_sph.cu
int gls_density(gl_ctx *c);
int gls_hydro(gl_ctx *c);
int gls_hydro_begin(gl_ctx *c);
int gls_hydro_launch(gl_ctx *c, const int *tgt, long long nt);
int gls_hydro_end(gl_ctx *c);
int gl_split_gas_targets(gl_ctx *c, long long split, long long *n_lo);
// gl_api.cu
int gl_prepare_targets(gl_ctx *c);
void gl_fresh_particle_set(gl_ctx *c);
// gl_let.cu (multi-rank: locally-essential tree exchange)
int gll_distributed_build(gl_ctx *c);	// what glt_build does when a communicator is attached
int gll_exchange_sph_state(gl_ctx *c);	// before force_update_hmax: Hsml + SphP of exported particles, hmax of exported cells
void gll_release(gl_ctx *c);
int glt_build_local(gl_ctx *c, bool entries);
int glt_launch_update_hmax(gl_ctx *c);
int gll_agree_violation(gl_ctx *c, int mine, int *any);
void gll_set_full(gl_ctx *c, bool full);
void gll_set_padding(gl_ctx *c, double fs);
bool gll_built_for(gl_ctx *c, bool bh, bool grav);
#define GL_LET_MAXRANKS 64	// the single-context build over the items currently loaded
#define GL_ERR_LET_VIOLATION 7001ull	// d_counter[2]: a walk wanted to open / a search touched a cut cell

#define GL_LAUNCH_COUNT(c) ((c)->kernel_launches++)

static inline int gl_grid(long long n, int block) { return (int) ((n + block - 1) / block); }

#ifdef __CUDACC__
// ---- split FP32 coordinates of the gravity walk ----
// q = (x - centre) * S lies in (-1, 1); hi = the nearest multiple of 2^-20 (exact in FP32), lo = the remainder
// (|lo| <= 2^-21, FP32 error <= 2^-46).  A coordinate difference (hi_a - hi_b) + (lo_a - lo_b) is then good to
// 2^-24 relative + 2^-43 absolute without any FP64 instruction in the walk.
#define GL_SPLIT_G 9.5367431640625e-07	// 2^-20
#define GL_SPLIT_GINV 1048576.0
__device__ __forceinline__ void gl_split(double x, double centre, double S, float &hi, float &lo)
{
  const double q = (x - centre) * S;
  const double k = rint(q * GL_SPLIT_GINV);
  const double H = k * GL_SPLIT_G;
  hi = (float) H;
  lo = (float) (q - H);
}
// read-only 32-byte record load as two 16-byte LDG.NC
__device__ __forceinline__ double4 gl_ld4(const double4 *p)
{
  const double2 *q = reinterpret_cast<const double2 *>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
#endif
