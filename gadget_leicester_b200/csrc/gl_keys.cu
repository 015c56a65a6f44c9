// gl_keys.cu - K1 domain extent, K2 Peano-Hilbert + Morton keys, K3 radix sort + SoA permutation.
//
// Reference behaviour reproduced (nothing is copied; the reference is CPU C):
//   domain_findExtent        domain.c:1972-2014   cube = 1.001 * max extent, fac = 2^21/len
//   peano_hilbert_key        peano.c:300-316      63-bit key, pix = x->4 | y->2 | z->1 per level
//   morton_key               peano.c:320-332      z->4, y->2, x->1 per level
//   peano_hilbert_order      peano.c:25-92        gas block and the rest sorted separately, stable
//
// The reference drives its Hilbert key from two 48x8 lookup tables.  Here the same curve is
// *generated*: the curve is defined by the order in which the 8 octants of a cell are visited
// (a Gray-code path) and, for every octant, the axis permutation + reflection that maps the
// parent's frame onto the child's.  Closing that definition under composition yields a 24-state
// automaton whose keys are identical to the reference's (checked bit-for-bit in tests/).
#include "gl_internal.h"
#include <cub/cub.cuh>

// ---------------------------------------------------------------- Hilbert automaton ----
namespace
{
struct Affine3
{				// T(p) = mask ^ XOR_b (bit b of p ? col[b] : 0), p a 3-bit octant code
  unsigned char col[3], mask;
};
inline unsigned apply(const Affine3 &T, unsigned p)
{
  unsigned v = T.mask;
  for(int b = 0; b < 3; b++)
    if((p >> b) & 1) v ^= T.col[b];
  return v;
}
inline Affine3 compose(const Affine3 &A, const Affine3 &B)	// A after B
{
  Affine3 R;
  for(int b = 0; b < 3; b++) R.col[b] = (unsigned char) (apply(A, B.col[b]) ^ A.mask);
  R.mask = (unsigned char) apply(A, B.mask);
  return R;
}
inline bool same(const Affine3 &A, const Affine3 &B)
{
  return A.col[0] == B.col[0] && A.col[1] == B.col[1] && A.col[2] == B.col[2] && A.mask == B.mask;
}

struct HilbertTables
{
  unsigned char sub[24][8];	// key digit emitted for octant code pix in state s
  unsigned char next[24][8];	// successor state
  int nstates;
};

// octant code = (x?4)|(y?2)|(z?1).  Visit order of the base cell (position along the curve):
//   (0,0,0) (0,1,0) (1,1,0) (1,0,0) (1,0,1) (1,1,1) (0,1,1) (0,0,1)
// rank_of[pix] is the curve position of octant pix.
const unsigned char kRankOf[8] = { 0, 7, 1, 6, 3, 4, 2, 5 };
// frame change applied when descending into the octant with (frame-local) code q:
// {image of bit0(z), bit1(y), bit2(x)}, reflection mask
const Affine3 kChild[8] = {
  {{2, 1, 4}, 0}, {{2, 1, 4}, 3}, {{4, 2, 1}, 0}, {{4, 2, 1}, 5},
  {{1, 2, 4}, 6}, {{1, 2, 4}, 6}, {{4, 2, 1}, 0}, {{4, 2, 1}, 5}
};

HilbertTables build_hilbert_tables()
{
  HilbertTables T;
  Affine3 states[64];
  int ns = 1;
  states[0] = Affine3{{1, 2, 4}, 0};
  for(int s = 0; s < ns; s++)
    for(unsigned pix = 0; pix < 8; pix++)
      {
	unsigned q = apply(states[s], pix);
	Affine3 nx = compose(kChild[q], states[s]);
	int id = -1;
	for(int k = 0; k < ns; k++)
	  if(same(states[k], nx)) id = k;
	if(id < 0)
	  {
	    id = ns;
	    states[ns++] = nx;
	  }
	T.sub[s][pix] = kRankOf[q];
	T.next[s][pix] = (unsigned char) id;
      }
  T.nstates = ns;		// 24 for this curve
  return T;
}

// built once; the initialisation of a function-local static is thread-safe (the scalar key function is called from the
// host pool's workers, gl_gadget_domain_keys)
const HilbertTables &hilbert_tables()
{
  static const HilbertTables T = build_hilbert_tables();
  return T;
}
}				// namespace

extern "C" gl_peanokey gl_peano_hilbert_key(int x, int y, int z, int bits)
{
  const HilbertTables &T = hilbert_tables();
  unsigned s = 0;
  gl_peanokey key = 0;
  for(int i = bits - 1; i >= 0; i--)
    {
      unsigned pix = (((x >> i) & 1) << 2) | (((y >> i) & 1) << 1) | ((z >> i) & 1);
      key = (key << 3) | T.sub[s][pix];
      s = T.next[s][pix];
    }
  return key;
}

extern "C" gl_peanokey gl_morton_key(int x, int y, int z, int bits)
{
  gl_peanokey m = 0;
  for(int i = bits - 1; i >= 0; i--)
    m = (m << 3) | ((((z >> i) & 1) << 2) | (((y >> i) & 1) << 1) | ((x >> i) & 1));
  return m;
}

__constant__ unsigned char c_hsub[24 * 8];
__constant__ unsigned char c_hnext[24 * 8];

static int upload_tables(gl_ctx *c)
{
  if(c->tables_up) return 0;	// per context: __constant__ memory is per device, and a process may open several GPUs
  const HilbertTables &T = hilbert_tables();
  if(T.nstates != 24) return c->fail("hilbert automaton closure != 24 states");
  GL_CUDA(c, cudaMemcpyToSymbol(c_hsub, T.sub, 24 * 8));
  GL_CUDA(c, cudaMemcpyToSymbol(c_hnext, T.next, 24 * 8));
  c->tables_up = true;
  return 0;
}

// ---------------------------------------------------------------- K1 extent ----
__global__ void k_extent_partial(const double *__restrict__ pos, long long n, double *__restrict__ part)
{
  double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
  for(long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    {
#pragma unroll
      for(int k = 0; k < 3; k++)
	{
	  double v = pos[3 * i + k];
	  lo[k] = fmin(lo[k], v);
	  hi[k] = fmax(hi[k], v);
	}
    }
  __shared__ double sh[GL_BLOCK / 32][6];
#pragma unroll
  for(int k = 0; k < 3; k++)
    for(int o = 16; o > 0; o >>= 1)
      {
	lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
	hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if(l == 0)
    for(int k = 0; k < 3; k++)
      {
	sh[w][k] = lo[k];
	sh[w][3 + k] = hi[k];
      }
  __syncthreads();
  if(threadIdx.x == 0)
    {
      for(int ww = 1; ww < blockDim.x / 32; ww++)
	for(int k = 0; k < 3; k++)
	  {
	    sh[0][k] = fmin(sh[0][k], sh[ww][k]);
	    sh[0][3 + k] = fmax(sh[0][3 + k], sh[ww][3 + k]);
	  }
      for(int k = 0; k < 6; k++) part[blockIdx.x * 6 + k] = sh[0][k];
    }
}

__global__ void k_extent_final(const double *__restrict__ part, int nblocks, double *__restrict__ out6)
{
  // single warp
  double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
  for(int b = threadIdx.x; b < nblocks; b += 32)
    for(int k = 0; k < 3; k++)
      {
	lo[k] = fmin(lo[k], part[b * 6 + k]);
	hi[k] = fmax(hi[k], part[b * 6 + 3 + k]);
      }
  for(int k = 0; k < 3; k++)
    for(int o = 16; o > 0; o >>= 1)
      {
	lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
	hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      }
  if(threadIdx.x == 0)
    for(int k = 0; k < 3; k++)
      {
	out6[k] = lo[k];
	out6[3 + k] = hi[k];
      }
}

static void domain_from_minmax(gl_domain &d, const double lo[3], const double hi[3])
{
  // domain.c:1999-2013, same operation order
  double len = 0;
  for(int j = 0; j < 3; j++)
    if(hi[j] - lo[j] > len) len = hi[j] - lo[j];
  len *= 1.001;
  for(int j = 0; j < 3; j++)
    {
      d.center[j] = 0.5 * (lo[j] + hi[j]);
      d.corner[j] = 0.5 * (lo[j] + hi[j]) - 0.5 * len;
    }
  d.len = len;
  d.fac = 1.0 / len * (double) (1ull << GL_BITS);
}

int glk_find_extent(gl_ctx *c)
{
  if(c->domain_imposed)
    {
      c->have_domain = true;
      return 0;
    }
  const int nb = 148 * 4;
  GL_CUDA(c, c->dom_dev.ensure((size_t) nb * 6 + 16));
  k_extent_partial<<<nb, GL_BLOCK, 0, c->stream>>>(c->in_pos.p, c->n, c->dom_dev.p + 16);
  GL_LAUNCH_COUNT(c);
  k_extent_final<<<1, 32, 0, c->stream>>>(c->dom_dev.p + 16, nb, c->dom_dev.p);
  GL_LAUNCH_COUNT(c);
  double mm[6];
  GL_CUDA(c, cudaMemcpyAsync(mm, c->dom_dev.p, sizeof(mm), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  domain_from_minmax(c->dom, mm, mm + 3);
  c->have_domain = true;
  return 0;
}

int glk_local_minmax(gl_ctx *c, double *dev6)
{
  const int nb = 148 * 4;
  GL_CUDA(c, c->dom_dev.ensure((size_t) nb * 6 + 16));
  k_extent_partial<<<nb, GL_BLOCK, 0, c->stream>>>(c->in_pos.p, c->n, c->dom_dev.p + 16);
  k_extent_final<<<1, 32, 0, c->stream>>>(c->dom_dev.p + 16, nb, dev6);
  c->kernel_launches += 2;
  return 0;
}

void glk_set_domain_minmax(gl_ctx *c, const double lo[3], const double hi[3])
{
  domain_from_minmax(c->dom, lo, hi);
  c->have_domain = true;
}

extern "C" int gl_domain_set_extent(gl_ctx *c, const double xmin[3], const double xmax[3])
{
  domain_from_minmax(c->dom, xmin, xmax);
  c->have_domain = true;
  c->domain_imposed = true;
  c->tree_valid = false;
  return 0;
}

// ---------------------------------------------------------------- K2 keys ----
__device__ __forceinline__ unsigned long long spread3(unsigned v)	// 21 bits -> every 3rd bit
{
  unsigned long long x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__device__ __forceinline__ void grid_coords(const double *__restrict__ pos, long long i, const gl_domain &d, int &ix,
					     int &iy, int &iz)
{
  // (int) ((P[i].Pos[k] - DomainCorner[k]) * DomainFac), forcetree.c:181-183 / domain.c - no FMA contraction
  ix = (int) __dmul_rn(__dsub_rn(pos[3 * i + 0], d.corner[0]), d.fac);
  iy = (int) __dmul_rn(__dsub_rn(pos[3 * i + 1], d.corner[1]), d.fac);
  iz = (int) __dmul_rn(__dsub_rn(pos[3 * i + 2], d.corner[2]), d.fac);
}

template <bool WITH_PH>
__global__ void k_keys(const double *__restrict__ pos, long long n, gl_domain d, unsigned long long *__restrict__ mkey,
		       unsigned long long *__restrict__ phkey, int *__restrict__ idx)
{
  __shared__ unsigned char s_sub[192], s_next[192];
  if(WITH_PH)
    {
      for(int t = threadIdx.x; t < 192; t += blockDim.x)
	{
	  s_sub[t] = c_hsub[t];
	  s_next[t] = c_hnext[t];
	}
      __syncthreads();
    }
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  int ix, iy, iz;
  grid_coords(pos, i, d, ix, iy, iz);
  mkey[i] = spread3((unsigned) ix) | (spread3((unsigned) iy) << 1) | (spread3((unsigned) iz) << 2);
  if(idx) idx[i] = (int) i;
  if(WITH_PH)
    {
      unsigned s = 0;
      unsigned long long key = 0;
#pragma unroll 1
      for(int b = GL_BITS - 1; b >= 0; b--)
	{
	  unsigned pix = (((ix >> b) & 1) << 2) | (((iy >> b) & 1) << 1) | ((iz >> b) & 1);
	  unsigned t = s * 8 + pix;
	  key = (key << 3) | s_sub[t];
	  s = s_next[t];
	}
      phkey[i] = key;
    }
}

// PH key from a Morton key: the Morton digit holds (z,y,x) where the Hilbert automaton wants (x,y,z)
__global__ void k_ph_from_morton(const unsigned long long *__restrict__ mkey, long long n,
				 unsigned long long *__restrict__ phkey, int *__restrict__ idx)
{
  __shared__ unsigned char s_sub[192], s_next[192];
  for(int t = threadIdx.x; t < 192; t += blockDim.x)
    {
      s_sub[t] = c_hsub[t];
      s_next[t] = c_hnext[t];
    }
  __syncthreads();
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const unsigned long long m = mkey[i];
  unsigned s = 0;
  unsigned long long key = 0;
#pragma unroll 1
  for(int b = GL_BITS - 1; b >= 0; b--)
    {
      const unsigned d = (unsigned) (m >> (3 * b)) & 7u;
      const unsigned pix = ((d & 1u) << 2) | (d & 2u) | (d >> 2);
      const unsigned t = s * 8 + pix;
      key = (key << 3) | s_sub[t];
      s = s_next[t];
    }
  phkey[i] = key;
  idx[i] = (int) i;
}

// ---------------------------------------------------------------- K3 sort + gather ----
__global__ void k_gather_sorted(long long nit, long long n, const int *__restrict__ perm, const double *__restrict__ pos,
				const double *__restrict__ vel, const double *__restrict__ mass,
				const double *__restrict__ hsml, const int *__restrict__ type,
				const double4 *__restrict__ c_sm, const double4 *__restrict__ c_aux, const signed char *__restrict__ c_level,
				double4 *__restrict__ s_pm, double4 *__restrict__ s_vh, int *__restrict__ s_type,
				signed char *__restrict__ s_lev, int *__restrict__ inv)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nit) return;
  int j = perm[i];
  if(j >= n)
    {
      // a cut cell of the locally-essential tree (multi-rank mode): an item that carries moments instead of a particle
      s_pm[i] = c_sm[j - n];
      s_vh[i] = c_aux[j - n];	// maxsoft, hmax, vmax: what a child cell contributes to the moments besides (s, mass)
      s_type[i] = -1;
      s_lev[i] = c_level[j - n];
      return;
    }
  s_pm[i] = make_double4(pos[3ll * j], pos[3ll * j + 1], pos[3ll * j + 2], mass[j]);
  s_vh[i] = make_double4(vel[3ll * j], vel[3ll * j + 1], vel[3ll * j + 2], hsml[j]);
  s_type[i] = type[j];
  s_lev[i] = GL_BITS + 1;
  inv[j] = (int) i;
}

// drop-in layer: gas velocities of the SPH sums are SphP[].VelPred (see gl_internal.h in_velpred)
__global__ void k_apply_velpred(long long ngas, const int *__restrict__ inv, const int *__restrict__ s_type,
				const double *__restrict__ velpred, double4 *__restrict__ s_vh)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= ngas) return;
  const int p = inv[i];
  if(s_type[p] != 0) return;
  s_vh[p].x = velpred[3 * i];
  s_vh[p].y = velpred[3 * i + 1];
  s_vh[p].z = velpred[3 * i + 2];
}

int glk_apply_velpred(gl_ctx *c)
{
  if(!c->have_velpred || c->n_velpred <= 0) return 0;
  k_apply_velpred<<<gl_grid(c->n_velpred, 256), 256, 0, c->stream>>>(c->n_velpred, c->inv.p, c->s_type.p, c->in_velpred.p, c->s_vh.p);
  GL_LAUNCH_COUNT(c);
  return 0;
}

static int sort_pairs(gl_ctx *c, const unsigned long long *kin, unsigned long long *kout, const int *vin, int *vout,
		      long long n)
{
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (int) n, 0, 63, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, kin, kout, vin, vout, (int) n, 0, 63, c->stream));
  c->kernel_launches += 5;	// onesweep: histogram + exclusive sum + ~3 digit passes (library-issued)
  return 0;
}

__global__ void k_iota(int *p, long long n, int base);

int glk_keys_and_sort(gl_ctx *c)
{
  // items of the sorted order: the n loaded particles and, in multi-rank mode, the ncell imported cut cells (sorted
  // in by the lowest Morton key of the cell - no particle of the set lies inside a cut cell)
  const long long n = c->n, nc = c->ncell, nit = n + nc;
  c->nit = nit;
  if(upload_tables(c)) return 1;
  if(!c->have_domain)
    if(glk_find_extent(c)) return 1;
  GL_CUDA(c, c->key_m.ensure(nit));
  GL_CUDA(c, c->key_m_sorted.ensure(nit));
  GL_CUDA(c, c->idx.ensure(nit));
  GL_CUDA(c, c->perm.ensure(nit));
  GL_CUDA(c, c->inv.ensure(n));
  GL_CUDA(c, c->s_pm.ensure(nit + 1));
  GL_CUDA(c, c->s_vh.ensure(nit + 1));
  GL_CUDA(c, c->s_type.ensure(nit + 1));
  GL_CUDA(c, c->s_lev.ensure(nit + 1));
  k_keys<false><<<gl_grid(n, 256), 256, 0, c->stream>>>(c->in_pos.p, n, c->dom, c->key_m.p, nullptr, c->idx.p);
  GL_LAUNCH_COUNT(c);
  if(nc > 0)
    {
      GL_CUDA(c, cudaMemcpyAsync(c->key_m.p + n, c->c_key.p, nc * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
      k_iota<<<gl_grid(nc, 256), 256, 0, c->stream>>>(c->idx.p + n, nc, (int) n);
      GL_LAUNCH_COUNT(c);
    }
  if(sort_pairs(c, c->key_m.p, c->key_m_sorted.p, c->idx.p, c->perm.p, nit)) return 1;
  k_gather_sorted<<<gl_grid(nit, 256), 256, 0, c->stream>>>(nit, n, c->perm.p, c->in_pos.p, c->in_vel.p, c->in_mass.p,
							    c->in_hsml.p, c->in_type.p, c->c_sm.p, c->c_aux.p, c->c_level.p, c->s_pm.p, c->s_vh.p,
							    c->s_type.p, c->s_lev.p, c->inv.p);
  GL_LAUNCH_COUNT(c);
  c->have_ph_order = false;
  // a caller that hands over its targets already ordered along the curve (a GPU's Peano-Hilbert segment, or the
  // reference's own P[] after peano_hilbert_order) does not need the second sort
  if(!(c->n_tgt_in >= 0 && c->tgt_ordered))
    if(glk_ph_sequence(c)) return 1;
  GL_CUDA(c, cudaGetLastError());
  return 0;
}

// Peano-Hilbert sequence of the tree-ordered particles: target groups (one warp = 32 consecutive
// targets) are formed along the Hilbert curve, which never jumps, so a group's bounding box stays
// compact; along the Morton curve the groups that straddle an octant boundary span the whole cell.
int glk_ph_sequence(gl_ctx *c)
{
  if(c->have_ph_order) return 0;
  if(c->ncell > 0) return c->fail("Peano-Hilbert sequence: not available on a locally-essential tree (multi-rank mode orders its own targets)");
  long long n = c->n;
  GL_CUDA(c, c->key_ph.ensure(n));
  GL_CUDA(c, c->ph_order.ensure(n + 1));
  k_ph_from_morton<<<gl_grid(n, 256), 256, 0, c->stream>>>(c->key_m_sorted.p, n, c->key_ph.p, c->idx.p);
  GL_LAUNCH_COUNT(c);
  if(sort_pairs(c, c->key_ph.p, c->key_m.p, c->idx.p, c->ph_order.p, n)) return 1;
  c->have_ph_order = true;
  return 0;
}

int glk_export_keys(gl_ctx *c, gl_peanokey *ph, gl_peanokey *mo, int where)
{
  long long n = c->n;
  if(upload_tables(c)) return 1;
  if(!c->have_domain)
    if(glk_find_extent(c)) return 1;
  GL_CUDA(c, c->key_m.ensure(n));
  GL_CUDA(c, c->key_ph.ensure(n));
  k_keys<true><<<gl_grid(n, 256), 256, 0, c->stream>>>(c->in_pos.p, n, c->dom, c->key_m.p, c->key_ph.p, nullptr);
  GL_LAUNCH_COUNT(c);
  cudaMemcpyKind kind = where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if(ph) GL_CUDA(c, cudaMemcpyAsync(ph, c->key_ph.p, n * sizeof(gl_peanokey), kind, c->stream));
  if(mo) GL_CUDA(c, cudaMemcpyAsync(mo, c->key_m.p, n * sizeof(gl_peanokey), kind, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tree_valid = false;	// key_m scratch is shared with the tree build
  return 0;
}

__global__ void k_iota(int *p, long long n, int base)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) p[i] = base + (int) i;
}

int glk_ph_order(gl_ctx *c, long long n_gas, int *order, int where)
{
  long long n = c->n;
  if(n_gas < 0 || n_gas > n) return c->fail("gl_peano_hilbert_order: bad n_gas");
  if(upload_tables(c)) return 1;
  if(!c->have_domain)
    if(glk_find_extent(c)) return 1;
  GL_CUDA(c, c->key_m.ensure(n));
  GL_CUDA(c, c->key_ph.ensure(n));
  GL_CUDA(c, c->key_m_sorted.ensure(n));
  GL_CUDA(c, c->idx.ensure(n));
  GL_CUDA(c, c->perm.ensure(n));
  k_keys<true><<<gl_grid(n, 256), 256, 0, c->stream>>>(c->in_pos.p, n, c->dom, c->key_m.p, c->key_ph.p, c->idx.p);
  GL_LAUNCH_COUNT(c);
  // peano.c:37-62 gas block, :64-88 the rest; mysort_peano is a stable merge sort, radix sort is stable too
  if(n_gas > 0)
    if(sort_pairs(c, c->key_ph.p, c->key_m_sorted.p, c->idx.p, c->perm.p, n_gas)) return 1;
  if(n - n_gas > 0)
    if(sort_pairs(c, c->key_ph.p + n_gas, c->key_m_sorted.p + n_gas, c->idx.p + n_gas, c->perm.p + n_gas, n - n_gas))
      return 1;
  cudaMemcpyKind kind = where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  GL_CUDA(c, cudaMemcpyAsync(order, c->perm.p, n * sizeof(int), kind, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tree_valid = false;
  return 0;
}

// ---------------------------------------------------------------- multi-rank helpers (gl_let.cu) ----
__device__ __forceinline__ unsigned long long ph_from_morton_dev(unsigned long long m, const unsigned char *s_sub, const unsigned char *s_next)
{
  unsigned s = 0;
  unsigned long long key = 0;
#pragma unroll 1
  for(int b = GL_BITS - 1; b >= 0; b--)
    {
      const unsigned d = (unsigned) (m >> (3 * b)) & 7u;
      const unsigned pix = ((d & 1u) << 2) | (d & 2u) | (d >> 2);
      const unsigned t = s * 8 + pix;
      key = (key << 3) | s_sub[t];
      s = s_next[t];
    }
  return key;
}

// Peano-Hilbert key hull of every stretch of the own particles (input order): block b reduces [bounds[b], bounds[b+1])
__global__ void k_ph_chunk_hulls(const unsigned long long *__restrict__ mkey_in, const long long *__restrict__ bounds,
				 unsigned long long *__restrict__ out)
{
  __shared__ unsigned char s_sub[192], s_next[192];
  __shared__ unsigned long long sh[8][2];
  for(int t = threadIdx.x; t < 192; t += blockDim.x)
    {
      s_sub[t] = c_hsub[t];
      s_next[t] = c_hnext[t];
    }
  __syncthreads();
  const int b = blockIdx.x;
  unsigned long long lo = ~0ull, hi = 0;
  for(long long i = bounds[b] + threadIdx.x; i < bounds[b + 1]; i += blockDim.x)
    {
      const unsigned long long k = ph_from_morton_dev(mkey_in[i], s_sub, s_next);
      lo = k < lo ? k : lo;
      hi = k > hi ? k : hi;
    }
  for(int o = 16; o > 0; o >>= 1)
    {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), bb = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = a < lo ? a : lo;
      hi = bb > hi ? bb : hi;
    }
  if((threadIdx.x & 31) == 0)
    {
      sh[threadIdx.x >> 5][0] = lo;
      sh[threadIdx.x >> 5][1] = hi;
    }
  __syncthreads();
  if(threadIdx.x == 0)
    {
      for(int w = 1; w < blockDim.x / 32; w++)
	{
	  lo = sh[w][0] < lo ? sh[w][0] : lo;
	  hi = sh[w][1] > hi ? sh[w][1] : hi;
	}
      out[2 * b] = lo;	// an empty stretch leaves lo > hi
      out[2 * b + 1] = hi;
    }
}

// key_m still holds the Morton keys in INPUT order after the sort (it was only read)
int glk_ph_chunk_hulls(gl_ctx *c, const long long *bounds_dev, int nchunks, unsigned long long *out)
{
  if(upload_tables(c)) return 1;
  k_ph_chunk_hulls<<<nchunks, 256, 0, c->stream>>>(c->key_m.p, bounds_dev, out);
  GL_LAUNCH_COUNT(c);
  return 0;
}

// ---- key hulls that do not depend on the caller's order -----------------------------------------------------------
// The stretch hulls above are tight while the caller's particles are in Peano-Hilbert order of the CURRENT grid.  When a
// rank keeps its particles over several steps (the resident loop, gl_resident.cu), the cube moves with the outermost
// particles, the grid shifts, and a few particles of every stretch land across a high-level cell boundary: their keys
// jump far away and the [min, max] hull of the stretch covers half the curve - nothing is missed (the hulls only have to
// COVER the rank's keys), but cells stop being "complete" and are exported particle by particle (measured: +18 ... +330 %
// imports).  Here the own keys are sorted and cut at the nchunks - 1 LARGEST GAPS: the smallest cover by nchunks
// intervals whatever the order of the particles - a stray particle becomes an interval of its own.
__global__ void k_ph_keys_of(long long n, const unsigned long long *__restrict__ mkey_in, unsigned long long *__restrict__ ph)
{
  __shared__ unsigned char s_sub[192], s_next[192];
  for(int t = threadIdx.x; t < 192; t += blockDim.x)
    {
      s_sub[t] = c_hsub[t];
      s_next[t] = c_hnext[t];
    }
  __syncthreads();
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) ph[i] = ph_from_morton_dev(mkey_in[i], s_sub, s_next);
}

__global__ void k_key_gaps(long long n, const unsigned long long *__restrict__ sorted, unsigned long long *__restrict__ gap, int *__restrict__ idx)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n - 1) return;
  gap[i] = sorted[i + 1] - sorted[i];
  idx[i] = (int) i;
}

// one block: cut[] = the positions of the ncut largest gaps (any order) -> ncut + 1 intervals of the sorted keys
__global__ void k_gap_hulls(long long n, int ncut, int nchunks, const unsigned long long *__restrict__ sorted, const int *__restrict__ cut,
			    unsigned long long *__restrict__ out)
{
  __shared__ int s_cut[256];
  const int t = threadIdx.x;
  if(t < ncut)
    {
      // rank of my cut among all cuts (ncut <= 255: a quadratic count is nothing)
      const int mine = cut[t];
      int r = 0;
      for(int k = 0; k < ncut; k++)
	{
	  const int o = cut[k];
	  r += (o < mine || (o == mine && k < t)) ? 1 : 0;
	}
      s_cut[r] = mine;
    }
  __syncthreads();
  if(t < nchunks)
    {
      if(t <= ncut)
	{
	  const long long a = t == 0 ? 0 : (long long) s_cut[t - 1] + 1, b = t == ncut ? n - 1 : (long long) s_cut[t];
	  out[2 * t] = sorted[a];
	  out[2 * t + 1] = sorted[b];
	}
      else
	{
	  out[2 * t] = ~0ull;	// unused hull: lo > hi
	  out[2 * t + 1] = 0;
	}
    }
}

int glk_ph_gap_hulls(gl_ctx *c, int nchunks, unsigned long long *out)
{
  const long long n = c->n_own;
  if(nchunks > 256 || n <= 0) return c->fail("key hulls: bad arguments");
  if(upload_tables(c)) return 1;
  GL_CUDA(c, c->hull_u64.ensure(3 * (size_t) n + 4));
  GL_CUDA(c, c->hull_i32.ensure(2 * (size_t) n + 4));
  unsigned long long *ph = c->hull_u64.p, *sorted = ph + n, *gap = sorted + n;
  int *idx = c->hull_i32.p, *idx_sorted = idx + n;
  k_ph_keys_of<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->key_m.p, ph);
  GL_LAUNCH_COUNT(c);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp, ph, sorted, (int) n, 0, 63, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tmp, ph, sorted, (int) n, 0, 63, c->stream));
  int ncut = (int) std::min<long long>(nchunks - 1, n - 1);
  if(ncut > 0)
    {
      k_key_gaps<<<gl_grid(n - 1, 256), 256, 0, c->stream>>>(n, sorted, gap, idx);
      GL_LAUNCH_COUNT(c);
      // (ph is free again: the sorted gaps go there)
      cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, gap, ph, idx, idx_sorted, (int) (n - 1), 0, 64, c->stream);
      GL_CUDA(c, c->cub_tmp.ensure(tmp));
      GL_CUDA(c, cub::DeviceRadixSort::SortPairsDescending(c->cub_tmp.p, tmp, gap, ph, idx, idx_sorted, (int) (n - 1), 0, 64, c->stream));
    }
  c->kernel_launches += 10;
  k_gap_hulls<<<1, 256, 0, c->stream>>>(n, ncut, nchunks, sorted, idx_sorted, out);
  GL_LAUNCH_COUNT(c);
  return 0;
}

// A cell of level L is a contiguous stretch of the Peano-Hilbert curve: all keys that share its leading 3 L bits.
__global__ void k_node_complete(long long m, const int *__restrict__ n_first, const signed char *__restrict__ n_level,
				const unsigned long long *__restrict__ mkey, const unsigned long long *__restrict__ ranges, int world,
				int nhull, long long stride, int me, unsigned char *__restrict__ complete)
{
  __shared__ unsigned char s_sub[192], s_next[192];
  for(int t = threadIdx.x; t < 192; t += blockDim.x)
    {
      s_sub[t] = c_hsub[t];
      s_next[t] = c_hnext[t];
    }
  __syncthreads();
  const long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(id >= m) return;
  const int L = n_level[id];
  const unsigned long long ph = ph_from_morton_dev(mkey[n_first[id]], s_sub, s_next);
  const int sh = 3 * (GL_BITS - L);
  const unsigned long long lo = (sh >= 63) ? 0ull : ((ph >> sh) << sh);
  const unsigned long long hi = (sh >= 63) ? ~0ull : (lo + ((1ull << sh) - 1ull));
  bool ok = true;
  for(int r = 0; r < world && ok; r++)
    {
      if(r == me) continue;
      const unsigned long long *hr = ranges + (long long) r * stride;
      for(int q = 0; q < nhull; q++)
	{
	  const unsigned long long a = hr[2 * q], b = hr[2 * q + 1];
	  if(a > b) continue;	// an empty stretch
	  if(!(b < lo || a > hi)) ok = false;
	}
    }
  complete[id] = ok ? 1 : 0;
}

int glk_node_complete(gl_ctx *c, const unsigned long long *ranges, int world, int nhull, long long stride, int me, unsigned char *complete)
{
  if(upload_tables(c)) return 1;
  if(c->nnodes <= 0) return 0;
  k_node_complete<<<gl_grid(c->nnodes, 256), 256, 0, c->stream>>>(c->nnodes, c->n_first.p, c->n_level.p, c->key_m_sorted.p, ranges, world,
								   nhull, stride, me, complete);
  GL_LAUNCH_COUNT(c);
  return 0;
}
