// gl_tree.cu - K4 bottom-up oct-tree from sorted Morton keys, K5 node moments, K9 hmax refresh.
//
// What the reference does (forcetree.c): force_treebuild_single (:125-373) inserts particles one by
// one into a top-down oct-tree on the 21-bit domain grid; force_update_node_recursive (:468-872)
// then computes, post-order and in child-slot order 0..7 (x->bit0, y->bit1, z->bit2), the FP64
// moments mass, s, hmax, vmax, maxsoft, and threads the tree (nextnode / sibling) for a stackless
// walk.  The set of nodes that can ever take part in an interaction is exactly "every oct-tree cell
// that holds >= 2 particles"; empty cells and single-particle cells of its always-complete top tree
// are opened without an interaction (forcetree.c:1996-2004) or carry zero mass.
//
// B200 design: particles are radix-sorted by Morton key (child-slot order == key order), so
//   c[i]  = number of oct-tree levels the keys i-1 and i share (c[0] = c[n] = -1)
// fully determines the tree.  A cell at level L whose first particle is a exists iff
// c[a] < L <= c[a+1]; so particle a opens max(0, c[a+1]-c[a]) nodes, one exclusive scan numbers all
// nodes, and that numbering IS the DFS pre-order the reference threads with nextnode/sibling:
//   node id = noff[a] + (L - c[a] - 1),  DFS entry index of the node = a + id,
//   DFS entry index of particle p = p + noff[p+1],  "sibling" of a node = entry index of its end.
// Moments are then evaluated level by level from the leaves up, each node summing its children in
// key order - the same FP64 operation order as the reference, hence bit-identical mass/s.
#include "gl_internal.h"
#include <cub/cub.cuh>
#include <climits>

__device__ __forceinline__ int common_level(unsigned long long a, unsigned long long b)
{
  unsigned long long x = a ^ b;
  if(x == 0) return GL_BITS;
  int p = 63 - __clzll((long long) x);	// highest differing bit, 0..62
  return (62 - p) / 3;
}

__global__ void k_levels(const unsigned long long *__restrict__ key, long long n, signed char *__restrict__ c)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i > n) return;
  if(i == 0 || i == n)
    c[i] = -1;
  else
    c[i] = (signed char) common_level(key[i - 1], key[i]);
}

// Items, not only particles: in multi-rank mode the sorted order also holds CUT CELLS (s_lev = level of the cell, 22 for
// a particle) - cells another rank has summarised for us.  A cut cell stands for >= 2 particles, so every proper ancestor
// of it exists as a node; the ones that start at item a reach down to level s_lev - 1 (chain nodes with one child).
__device__ __forceinline__ int top_level_at(const signed char *__restrict__ c, const signed char *__restrict__ s_lev, long long a)
{
  const int up = c[a + 1], lv = s_lev[a];
  return (lv <= GL_BITS && lv - 1 > up) ? lv - 1 : up;
}

__global__ void k_count(const signed char *__restrict__ c, const signed char *__restrict__ s_lev, long long n, int *__restrict__ cnt)
{
  long long a = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(a > n) return;
  int v = 0;
  if(a < n)
    {
      v = top_level_at(c, s_lev, a) - (int) c[a];
      if(v < 0) v = 0;
    }
  cnt[a] = v;
}

__global__ void k_make_nodes(const signed char *__restrict__ c, const signed char *__restrict__ s_lev, const int *__restrict__ noff, long long n,
			     int *__restrict__ n_first, signed char *__restrict__ n_level, int *__restrict__ levcount)
{
  __shared__ int sh[GL_BITS + 1];
  if(threadIdx.x <= GL_BITS) sh[threadIdx.x] = 0;
  __syncthreads();
  long long a = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(a < n)
    {
      int c0 = c[a], c1 = top_level_at(c, s_lev, a);
      int base = noff[a];
      for(int L = c0 + 1; L <= c1; L++)
	{
	  int id = base + (L - c0 - 1);
	  n_first[id] = (int) a;
	  n_level[id] = (signed char) L;
	  atomicAdd(&sh[L], 1);
	}
    }
  __syncthreads();
  if(threadIdx.x <= GL_BITS && sh[threadIdx.x]) atomicAdd(&levcount[threadIdx.x], sh[threadIdx.x]);
}

__global__ void k_fill_levels(const signed char *__restrict__ n_level, long long m, int *__restrict__ cursor,
			      int *__restrict__ lev_list)
{
  long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  bool ok = id < m;
  int L = ok ? n_level[id] : 31;
  unsigned peers = __match_any_sync(0xffffffffu, L);
  int lane = threadIdx.x & 31;
  int leader = __ffs(peers) - 1;
  int rank = __popc(peers & ((1u << lane) - 1));
  int base = 0;
  if(ok && lane == leader) base = atomicAdd(&cursor[L], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if(ok) lev_list[base + rank] = (int) id;
}

// first index j in (lo, hi] such that (key[j] >> sh) != pref, or hi if none; galloping then bisection
__device__ __forceinline__ int range_end(const unsigned long long *__restrict__ key, int lo, int hi, int sh,
					 unsigned long long pref)
{
  int step = 1, good = lo;	// key[good] matches
  int bad = hi;			// first known non-match (or hi)
  while(true)
    {
      long long probe = (long long) good + step;
      if(probe >= hi) break;
      if((key[probe] >> sh) == pref)
	{
	  good = (int) probe;
	  step <<= 1;
	}
      else
	{
	  bad = (int) probe;
	  break;
	}
    }
  while(bad - good > 1)
    {
      int mid = good + (bad - good) / 2;
      if((key[mid] >> sh) == pref)
	good = mid;
      else
	bad = mid;
    }
  return bad;
}

__global__ void k_moments(int L, const int *__restrict__ lev_list, int count, const unsigned long long *__restrict__ key,
			  long long n, const signed char *__restrict__ c, const int *__restrict__ noff,
			  const int *__restrict__ n_first, int *__restrict__ n_end, int *__restrict__ n_parent,
			  int *__restrict__ p_father, const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh,
			  const int *__restrict__ s_type, const signed char *__restrict__ s_lev, double4 *__restrict__ n_sm,
			  double4 *__restrict__ n_cl, double4 *__restrict__ n_aux, int *__restrict__ n_child, gl_domain dom, gl_dparams par)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= count) return;
  bool cut = false;		// a cut cell somewhere below this node (multi-rank mode)
  int id = lev_list[t];
  // child table for the parallel SPH walks: DFS entry index of a child node, -(p+2) for a particle leaf p,
  // -1 = no further child; a finest-level bucket (coincident keys) is flagged by n_child[0] = INT_MIN
  int nch = 0;
  int *child_tab = n_child + 8ll * id;
  int a = n_first[id];
  unsigned long long ka = key[a];
  if(L == 0) n_parent[id] = -1;

  // geometric centre and side length: root = domain cube, then the +-len/4 steps of forcetree.c:259-277 / :404-411
  double len = dom.len, cx = dom.center[0], cy = dom.center[1], cz = dom.center[2];
  for(int lev = 1; lev <= L; lev++)
    {
      int d = (int) ((ka >> (3 * (GL_BITS - lev))) & 7);
      double lenhalf = __dmul_rn(0.25, len);
      cx = (d & 1) ? __dadd_rn(cx, lenhalf) : __dsub_rn(cx, lenhalf);
      cy = (d & 2) ? __dadd_rn(cy, lenhalf) : __dsub_rn(cy, lenhalf);
      cz = (d & 4) ? __dadd_rn(cz, lenhalf) : __dsub_rn(cz, lenhalf);
      len = __dmul_rn(0.5, len);
    }

  // children in slot order (forcetree.c:543-729).  Child and node boundaries are read off c[] (one BYTE per particle:
  // c[i] = number of levels keys i-1 and i share): inside the node c[i] > L means "same child", c[i] == L "next
  // child", c[i] < L "node ends".  The galloping searches on the 8-byte keys this replaces cost ~15 scattered
  // sectors per node; cells with more than 256 particles fall back to them for the rest of their range.
  double mass = 0, s0 = 0, s1 = 0, s2 = 0, hmax = 0, vmax = 0, maxsoft = 0;
  int p = a, scanned = 0, b;
  const int sh2 = 3 * (GL_BITS - 1 - L);
  while(true)
    {
      int q = p + 1;
      if(L < GL_BITS)
	{
	  while(scanned < 256 && c[q] > L)
	    {
	      q++;
	      scanned++;
	    }
	  if(c[q] > L) q = range_end(key, q - 1, (int) n, sh2, key[p] >> sh2);	// big cell: finish this child by bisection
	}
      const int lvp = s_lev[p];
      if(q - p == 1 && lvp <= GL_BITS && lvp <= L + 1)
	{
	  // cut-cell leaf: the moments its owner computed, added exactly like a child cell's (s_vh = maxsoft, hmax, vmax)
	  const double4 cs = s_pm[p], ca = s_vh[p];
	  mass = __dadd_rn(mass, cs.w);
	  s0 = __dadd_rn(s0, __dmul_rn(cs.w, cs.x));
	  s1 = __dadd_rn(s1, __dmul_rn(cs.w, cs.y));
	  s2 = __dadd_rn(s2, __dmul_rn(cs.w, cs.z));
	  if(ca.y > hmax) hmax = ca.y;
	  if(ca.z > vmax) vmax = ca.z;
	  if(ca.x > maxsoft) maxsoft = ca.x;
	  p_father[p] = id;
	  if(nch < 8) child_tab[nch] = p + noff[p + 1];	// its DFS entry (a node-type record without descendants)
	  nch++;
	  cut = true;
	}
      else if(q - p == 1 && lvp > GL_BITS)
	{
	  // particle leaf (for L == GL_BITS every particle of the finest cell hangs directly below)
	  double4 pm = s_pm[p];
	  double4 vh = s_vh[p];
	  int ty = s_type[p];
	  mass = __dadd_rn(mass, pm.w);
	  s0 = __dadd_rn(s0, __dmul_rn(pm.w, pm.x));
	  s1 = __dadd_rn(s1, __dmul_rn(pm.w, pm.y));
	  s2 = __dadd_rn(s2, __dmul_rn(pm.w, pm.z));
	  if(ty == 0)
	    {
	      if(vh.w > hmax) hmax = vh.w;
	      if(vh.w > maxsoft) maxsoft = vh.w;
	    }
	  else if(par.soft[ty] > maxsoft)
	    maxsoft = par.soft[ty];
	  vmax = fmax(vmax, fmax(fabs(vh.x), fmax(fabs(vh.y), fabs(vh.z))));
	  p_father[p] = id;
	  if(nch < 8) child_tab[nch] = -(p + 2);
	  nch++;
	}
      else
	{
	  int child = noff[p] + (L - (int) c[p]);
	  double4 cs = n_sm[child];
	  double4 ca = n_aux[child];
	  mass = __dadd_rn(mass, cs.w);
	  s0 = __dadd_rn(s0, __dmul_rn(cs.w, cs.x));
	  s1 = __dadd_rn(s1, __dmul_rn(cs.w, cs.y));
	  s2 = __dadd_rn(s2, __dmul_rn(cs.w, cs.z));
	  if(ca.y > hmax) hmax = ca.y;
	  if(ca.z > vmax) vmax = ca.z;
	  if(ca.x > maxsoft) maxsoft = ca.x;
	  if(ca.w != 0) cut = true;
	  n_parent[child] = id;
	  if(nch < 8) child_tab[nch] = p + child;	// DFS entry index of node (p, L+1): first particle + node id
	  nch++;
	}
      if(c[q] < L)
	{
	  b = q;		// keys q-1 and q part company above this cell's level (c[n] = -1 ends the last cell)
	  break;
	}
      p = q;
      scanned++;
    }
  n_end[id] = b;
  if(mass != 0)
    {
      s0 = __ddiv_rn(s0, mass);
      s1 = __ddiv_rn(s1, mass);
      s2 = __ddiv_rn(s2, mass);
    }
  else
    {
      s0 = cx;
      s1 = cy;
      s2 = cz;
    }
  n_sm[id] = make_double4(s0, s1, s2, mass);
  n_cl[id] = make_double4(cx, cy, cz, len);
  n_aux[id] = make_double4(maxsoft, hmax, vmax, cut ? 1.0 : 0.0);
  for(int k = nch; k < 8; k++) child_tab[k] = -1;
  if(nch > 8) child_tab[0] = INT_MIN;
}

// first index in [lo, hi) whose (key >> sh) is >= pref (keys are sorted), plain bisection
__device__ __forceinline__ int lower_bound_prefix(const unsigned long long *__restrict__ key, int lo, int hi, int sh,
						  unsigned long long pref)
{
  while(lo < hi)
    {
      const int mid = lo + (hi - lo) / 2;
      if((key[mid] >> sh) < pref)
	lo = mid + 1;
      else
	hi = mid;
    }
  return lo;
}

// k_moments for the sparsely populated upper levels: EIGHT LANES PER NODE.  A cell near the root spans 10^5..10^6
// particles, and finding its eight child ranges one after the other is a chain of ~300 dependent loads - with one
// thread per node the six top levels of the 1.6M disc took 0.7 of the 1.07 ms of all moment launches.  Here lane s
// bisects for the start of octant s and fetches that child's record, all eight in parallel; the sums are then formed
// in slot order with exactly the operations of k_moments (bit-identical moments, forcetree.c:543-729).
__global__ void k_moments_wide(int L, const int *__restrict__ lev_list, int count, const unsigned long long *__restrict__ key,
			       long long n, const signed char *__restrict__ c, const int *__restrict__ noff,
			       const int *__restrict__ n_first, int *__restrict__ n_end, int *__restrict__ n_parent,
			       int *__restrict__ p_father, const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh,
			       const int *__restrict__ s_type, const signed char *__restrict__ s_lev, double4 *__restrict__ n_sm,
			       double4 *__restrict__ n_cl, double4 *__restrict__ n_aux, int *__restrict__ n_child, gl_domain dom, gl_dparams par)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = t >> 3, s = t & 7;
  const int lane = threadIdx.x & 31, lbase = lane & ~7;
  const bool live = g < count;
  const int id = lev_list[live ? g : 0];
  const int a = n_first[id];
  const unsigned long long ka = key[a];
  const int sh = 3 * (GL_BITS - L);	// the node's own prefix
  const int sh2 = sh - 3;		// its children's
  // end of the node's range: first key beyond the node's prefix (lane 0 of the group would do; all eight agree)
  const int b = (L == 0) ? (int) n : lower_bound_prefix(key, a, (int) n, sh, (ka >> sh) + 1);
  double len = dom.len, cx = dom.center[0], cy = dom.center[1], cz = dom.center[2];
  for(int lev = 1; lev <= L; lev++)
    {
      int d = (int) ((ka >> (3 * (GL_BITS - lev))) & 7);
      double lenhalf = __dmul_rn(0.25, len);
      cx = (d & 1) ? __dadd_rn(cx, lenhalf) : __dsub_rn(cx, lenhalf);
      cy = (d & 2) ? __dadd_rn(cy, lenhalf) : __dsub_rn(cy, lenhalf);
      cz = (d & 4) ? __dadd_rn(cz, lenhalf) : __dsub_rn(cz, lenhalf);
      len = __dmul_rn(0.5, len);
    }
  // child range of octant s: [lo, hi)
  const unsigned long long cpref = ((ka >> sh) << 3) | (unsigned long long) s;
  const int lo = lower_bound_prefix(key, a, b, sh2, cpref);
  int hi = __shfl_sync(0xffffffffu, lo, lbase + ((s + 1) & 7));
  if(s == 7) hi = b;
  const int np = hi - lo;
  // this lane's child: mass, m*x, m*y, m*z, hmax, vmax, maxsoft
  double cm = 0, t0 = 0, t1 = 0, t2 = 0, chm = 0, cvm = 0, cms = 0;
  int code = -1;
  bool mycut = false;
  const int lvlo = (live && np == 1) ? (int) s_lev[lo] : GL_BITS + 1;
  if(live && np == 1 && lvlo <= GL_BITS && lvlo <= L + 1)
    {
      // cut-cell leaf (multi-rank mode): its owner's moments, s_vh = (maxsoft, hmax, vmax)
      const double4 cs = s_pm[lo], ca = s_vh[lo];
      cm = cs.w;
      t0 = __dmul_rn(cs.w, cs.x); t1 = __dmul_rn(cs.w, cs.y); t2 = __dmul_rn(cs.w, cs.z);
      chm = ca.y; cvm = ca.z; cms = ca.x;
      p_father[lo] = id;
      code = lo + noff[lo + 1];
      mycut = true;
    }
  else if(live && np == 1 && lvlo > GL_BITS)
    {
      const double4 pm = s_pm[lo];
      const double4 vh = s_vh[lo];
      const int ty = s_type[lo];
      cm = pm.w;
      t0 = __dmul_rn(pm.w, pm.x); t1 = __dmul_rn(pm.w, pm.y); t2 = __dmul_rn(pm.w, pm.z);
      if(ty == 0)
	{
	  chm = vh.w;
	  cms = vh.w;
	}
      else
	cms = par.soft[ty];
      cvm = fmax(fabs(vh.x), fmax(fabs(vh.y), fabs(vh.z)));
      p_father[lo] = id;
      code = -(lo + 2);
    }
  else if(live && np >= 1)	// several items, or one cut cell that lies deeper than the next level (chain of one-child cells)
    {
      const int child = noff[lo] + (L - (int) c[lo]);
      const double4 cs = n_sm[child];
      const double4 ca = n_aux[child];
      cm = cs.w;
      t0 = __dmul_rn(cs.w, cs.x); t1 = __dmul_rn(cs.w, cs.y); t2 = __dmul_rn(cs.w, cs.z);
      chm = ca.y; cvm = ca.z; cms = ca.x;
      mycut = ca.w != 0;
      n_parent[child] = id;
      code = lo + child;	// DFS entry index of the child node
    }
  const bool present = live && np > 0;
  const unsigned pres8 = (__ballot_sync(0xffffffffu, present) >> lbase) & 0xffu;
  const bool cut = ((__ballot_sync(0xffffffffu, mycut) >> lbase) & 0xffu) != 0;
  // slot-ordered sums, every lane of the group forms the same values
  double mass = 0, s0 = 0, s1 = 0, s2 = 0, hmax = 0, vmax = 0, maxsoft = 0;
#pragma unroll
  for(int k = 0; k < 8; k++)
    {
      const double m_k = __shfl_sync(0xffffffffu, cm, lbase + k), a0 = __shfl_sync(0xffffffffu, t0, lbase + k),
	a1 = __shfl_sync(0xffffffffu, t1, lbase + k), a2 = __shfl_sync(0xffffffffu, t2, lbase + k);
      const double h_k = __shfl_sync(0xffffffffu, chm, lbase + k), v_k = __shfl_sync(0xffffffffu, cvm, lbase + k),
	ms_k = __shfl_sync(0xffffffffu, cms, lbase + k);
      if((pres8 >> k) & 1)
	{
	  mass = __dadd_rn(mass, m_k);
	  s0 = __dadd_rn(s0, a0); s1 = __dadd_rn(s1, a1); s2 = __dadd_rn(s2, a2);
	  if(h_k > hmax) hmax = h_k;
	  if(v_k > vmax) vmax = v_k;
	  if(ms_k > maxsoft) maxsoft = ms_k;
	}
    }
  if(!live) return;
  // child table: present children in slot order, then -1
  int *child_tab = n_child + 8ll * id;
  const int pos = __popc(pres8 & ((1u << s) - 1));
  const int nch = __popc(pres8);
  if(present) child_tab[pos] = code;
  if(s >= nch) child_tab[s] = -1;
  if(s == 0)
    {
      n_end[id] = b;
      if(L == 0) n_parent[id] = -1;
      if(mass != 0)
	{
	  s0 = __ddiv_rn(s0, mass);
	  s1 = __ddiv_rn(s1, mass);
	  s2 = __ddiv_rn(s2, mass);
	}
      else
	{
	  s0 = cx;
	  s1 = cy;
	  s2 = cz;
	}
      n_sm[id] = make_double4(s0, s1, s2, mass);
      n_cl[id] = make_double4(cx, cy, cz, len);
      n_aux[id] = make_double4(maxsoft, hmax, vmax, cut ? 1.0 : 0.0);
    }
}

// Scales of the FP32 walk records: S = 2^-floor(log2(DomainLen)) makes the cube [1,2) wide, Ms = 2^-floor(log2(total
// mass)) the root mass [1,2).  Powers of two, so scaled and unscaled comparisons decide alike; FP32 then never
// overflows or underflows whatever the unit system.
__global__ void k_grav_scale(long long m, long long n, const double4 *__restrict__ n_sm, const double4 *__restrict__ s_pm,
			     double domlen, double *__restrict__ scale)
{
  if(blockIdx.x || threadIdx.x) return;
  double mt = (m > 0) ? n_sm[0].w : s_pm[0].w;
  if(!(mt > 0) || !isfinite(mt)) mt = 1.0;
  scale[0] = (domlen > 0 && isfinite(domlen)) ? exp2(-floor(log2(domlen))) : 1.0;	// one particle: a cube of size 0
  scale[1] = exp2(-floor(log2(mt)));
}

__global__ void k_entries_particles(long long n, const int *__restrict__ noff, const double4 *__restrict__ s_pm,
				    const double4 *__restrict__ s_vh, const int *__restrict__ s_type,
				    const signed char *__restrict__ s_lev, const unsigned long long *__restrict__ key,
				    int *__restrict__ e_next, double4 *__restrict__ e_A, double4 *__restrict__ e_B,
				    double2 *__restrict__ e_C, double *__restrict__ e_hmax, int2 *__restrict__ e_rng, float4 *__restrict__ e_F0,
				    float4 *__restrict__ e_F1, float4 *__restrict__ e_F2, const double *__restrict__ scale, gl_domain dom,
				    gl_dparams par)
{
  long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(p >= n) return;
  long long e = p + noff[p + 1];
  const double S = scale[0], Ms = scale[1];
  const double4 pm = s_pm[p];
  float4 f0, f1;
  gl_split(pm.x, dom.center[0], S, f0.x, f1.x);
  gl_split(pm.y, dom.center[1], S, f0.y, f1.y);
  gl_split(pm.z, dom.center[2], S, f0.z, f1.z);
  f0.w = (float) (pm.w * Ms);
  const int lv = s_lev[p];
  if(lv <= GL_BITS)
    {
      // CUT CELL (multi-rank mode): a cell record whose descendants stayed with their owner.  It takes part like any
      // accepted cell; a walk that wants to open it, or a neighbour search that touches it, raises
      // GL_ERR_LET_VIOLATION instead (e_next = e + 1 and e_rng = (-1,-1) mark it).  Centre and side length by the
      // same +-len/4 steps as k_moments, so they are the values of the global tree.
      const unsigned long long ka = key[p];
      double len = dom.len, cx = dom.center[0], cy = dom.center[1], cz = dom.center[2];
      for(int l = 1; l <= lv; l++)
	{
	  int d = (int) ((ka >> (3 * (GL_BITS - l))) & 7);
	  double lenhalf = __dmul_rn(0.25, len);
	  cx = (d & 1) ? __dadd_rn(cx, lenhalf) : __dsub_rn(cx, lenhalf);
	  cy = (d & 2) ? __dadd_rn(cy, lenhalf) : __dsub_rn(cy, lenhalf);
	  cz = (d & 4) ? __dadd_rn(cz, lenhalf) : __dsub_rn(cz, lenhalf);
	  len = __dmul_rn(0.5, len);
	}
      const double4 ax = s_vh[p];	// maxsoft, hmax, vmax
      e_next[e] = (int) e + 1;
      e_A[e] = pm;
      e_B[e] = make_double4(cx, cy, cz, len);
      e_C[e] = make_double2(ax.x, __dmul_rn(__dmul_rn(pm.w, len), len));
      e_hmax[e] = ax.y;
      e_rng[e] = make_int2(-1, -1);
      f1.w = (float) (len * S);
      e_F0[e] = f0;
      e_F1[e] = f1;
      e_F2[e] = make_float4((float) ((cx - pm.x) * S), (float) ((cy - pm.y) * S), (float) ((cz - pm.z) * S), (float) (ax.x * S));
      return;
    }
  int ty = s_type[p];
  double soft = (ty == 0) ? s_vh[p].w : par.soft[ty];
  e_next[e] = -1 - (int) (p * 8 + ty);
  e_A[e] = pm;
  e_B[e] = make_double4(soft, 0.0, 0.0, -1.0);
  e_C[e] = make_double2(0.0, 0.0);
  e_hmax[e] = 0.0;
  f1.w = (float) (soft * S);
  e_F0[e] = f0;
  e_F1[e] = f1;
}

__global__ void k_entries_nodes(long long m, long long n, const int *__restrict__ noff, const int *__restrict__ n_first,
				const int *__restrict__ n_end, const double4 *__restrict__ n_sm,
				const double4 *__restrict__ n_cl, const double4 *__restrict__ n_aux,
				int *__restrict__ e_next, double4 *__restrict__ e_A, double4 *__restrict__ e_B,
				double2 *__restrict__ e_C, double *__restrict__ e_hmax, float4 *__restrict__ e_F0,
				float4 *__restrict__ e_F1, float4 *__restrict__ e_F2, const double *__restrict__ scale,
				gl_domain dom)
{
  long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(id >= m) return;
  long long e = (long long) n_first[id] + id;
  int b = n_end[id];
  e_next[e] = b + noff[b];	// "sibling": first DFS entry behind this subtree
  double4 sm = n_sm[id];
  double4 cl = n_cl[id];
  double4 ax = n_aux[id];
  e_A[e] = sm;
  e_B[e] = cl;
  // mass * nop->len * nop->len, evaluated left to right as in forcetree.c:2085
  e_C[e] = make_double2(ax.x, __dmul_rn(__dmul_rn(sm.w, cl.w), cl.w));
  e_hmax[e] = ax.y;
  const double S = scale[0], Ms = scale[1];
  float4 f0, f1;
  gl_split(sm.x, dom.center[0], S, f0.x, f1.x);
  gl_split(sm.y, dom.center[1], S, f0.y, f1.y);
  gl_split(sm.z, dom.center[2], S, f0.z, f1.z);
  f0.w = (float) (sm.w * Ms);
  f1.w = (float) (cl.w * S);
  e_F0[e] = f0;
  e_F1[e] = f1;
  e_F2[e] = make_float4((float) ((cl.x - sm.x) * S), (float) ((cl.y - sm.y) * S), (float) ((cl.z - sm.z) * S), (float) (ax.x * S));
}

// particle range [a,b) below every node entry: lets the SPH walks swallow a small cell in one coalesced step
__global__ void k_entries_range(long long m, const int *__restrict__ n_first, const int *__restrict__ n_end,
				const double4 *__restrict__ n_aux, int2 *__restrict__ e_rng)
{
  long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(id >= m) return;
  // node id = entry index - first.  A cell with a cut cell somewhere below is never swallowed whole by the SPH
  // searches (end = INT_MAX makes it "large"): they descend, so the cut cell itself meets the cube test.
  e_rng[(long long) n_first[id] + id] = make_int2(n_first[id], n_aux[id].w != 0 ? INT_MAX : n_end[id]);
}

static int scan_exclusive(gl_ctx *c, const int *in, int *out, long long n)
{
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int) n, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, in, out, (int) n, c->stream));
  c->kernel_launches += 2;
  return 0;
}

gl_dparams gl_make_dparams(const gl_ctx *c)
{
  gl_dparams d{};
  d.theta = c->par.ErrTolTheta;
  d.errtolforceacc = c->par.ErrTolForceAcc;
  for(int i = 0; i < 6; i++) d.soft[i] = c->par.ForceSoftening[i];
  for(int i = 0; i < 6; i++) d.softtab[i] = c->par.SofteningTable[i];
  d.G = c->par.G;
  d.lambda_fac = c->par.OmegaLambdaHubble2;
  d.desnumngb = c->par.DesNumNgb;
  d.maxdev = c->par.MaxNumNgbDeviation;
  d.bhfac = c->par.BlackHoleNgbFactor;
  d.mingashsml = c->par.MinGasHsml;
  d.artvisc = c->par.ArtBulkViscConst;
  d.timebase_interval = c->par.Timebase_interval;
  d.ti_current = c->par.Ti_Current;
  return d;
}

int glt_build(gl_ctx *c)
{
  // multi-rank mode: keys of the global cube, own tree, locally-essential exchange, merged tree (gl_let.cu)
  if(c->comm && c->comm->world > 1) return gll_distributed_build(c);
  c->n_own = c->n;
  c->ncell = 0;
  return glt_build_local(c, true);
}

// The build over the items currently loaded: c->n particles (+ c->ncell cut cells).  entries = false stops after the
// moments (the own-particle tree a rank only selects its exports from).
int glt_build_local(gl_ctx *c, bool entries)
{
  if(c->n <= 0) return c->fail("force_treebuild: no particles loaded");
  if(!c->have_par) return c->fail("force_treebuild: gl_set_params not called");
  cudaEventRecord(c->ev[0], c->stream);
  if(glk_keys_and_sort(c)) return 1;
  const long long n = c->nit;	// items of the sorted order
  cudaEventRecord(c->ev[1], c->stream);
  const unsigned long long *key = c->key_m_sorted.p;
  GL_CUDA(c, c->lev_c.ensure(n + 2));
  GL_CUDA(c, c->ncnt.ensure(n + 2));
  GL_CUDA(c, c->noff.ensure(n + 2));
  GL_CUDA(c, c->scan_tmp_i.ensure(64));
  k_levels<<<gl_grid(n + 1, 256), 256, 0, c->stream>>>(key, n, c->lev_c.p);
  k_count<<<gl_grid(n + 1, 256), 256, 0, c->stream>>>(c->lev_c.p, c->s_lev.p, n, c->ncnt.p);
  c->kernel_launches += 2;
  if(scan_exclusive(c, c->ncnt.p, c->noff.p, n + 1)) return 1;
  int m = 0;
  GL_CUDA(c, cudaMemcpyAsync(&m, c->noff.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->nnodes = m;
  c->nentries = n + m;
  if(c->nentries >= (1ll << 28)) return c->fail("force_treebuild: more than 2^28 tree entries (index encoding limit)");
  size_t mm = (size_t) m + 1;
  GL_CUDA(c, c->n_first.ensure(mm));
  GL_CUDA(c, c->n_end.ensure(mm));
  GL_CUDA(c, c->n_parent.ensure(mm));
  GL_CUDA(c, c->n_level.ensure(mm));
  GL_CUDA(c, c->lev_list.ensure(mm));
  GL_CUDA(c, c->n_sm.ensure(mm));
  GL_CUDA(c, c->n_cl.ensure(mm));
  GL_CUDA(c, c->n_aux.ensure(mm));
  GL_CUDA(c, c->n_child.ensure(8 * mm));
  GL_CUDA(c, c->p_father.ensure(n + 1));
  int *levcount = c->scan_tmp_i.p;	// [0..21] counts, [32..53] cursors
  GL_CUDA(c, cudaMemsetAsync(levcount, 0, 64 * sizeof(int), c->stream));
  GL_CUDA(c, cudaMemsetAsync(c->p_father.p, 0xff, (n + 1) * sizeof(int), c->stream));
  int hcount[GL_BITS + 1];
  if(m > 0)
    {
      k_make_nodes<<<gl_grid(n, 256), 256, 0, c->stream>>>(c->lev_c.p, c->s_lev.p, c->noff.p, n, c->n_first.p, c->n_level.p,
							  levcount);
      GL_LAUNCH_COUNT(c);
      GL_CUDA(c, cudaMemcpyAsync(hcount, levcount, sizeof(hcount), cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      int start = 0, hcur[GL_BITS + 1];
      for(int L = 0; L <= GL_BITS; L++)
	{
	  c->lev_start[L] = start;
	  hcur[L] = start;
	  start += hcount[L];
	}
      c->lev_start[GL_BITS + 1] = start;
      if(start != m) return c->fail("force_treebuild: level histogram does not add up");
      GL_CUDA(c, cudaMemcpyAsync(levcount + 32, hcur, sizeof(hcur), cudaMemcpyHostToDevice, c->stream));
      k_fill_levels<<<gl_grid(m, 256), 256, 0, c->stream>>>(c->n_level.p, m, levcount + 32, c->lev_list.p);
      GL_LAUNCH_COUNT(c);
    }
  else
    for(int L = 0; L <= GL_BITS + 1; L++) c->lev_start[L] = 0;
  cudaEventRecord(c->ev[2], c->stream);
  gl_dparams dp = gl_make_dparams(c);
  if(m > 0)
    for(int L = GL_BITS; L >= 0; L--)
      {
	int cnt = c->lev_start[L + 1] - c->lev_start[L];
	if(cnt <= 0) continue;
	static const int wide_max = getenv("GL_WIDE_MAX") ? atoi(getenv("GL_WIDE_MAX")) : 65536;
	if(L < GL_BITS && cnt <= wide_max)	// few, fat cells: eight lanes per node (see k_moments_wide)
	  k_moments_wide<<<gl_grid(8ll * cnt, 128), 128, 0, c->stream>>>(L, c->lev_list.p + c->lev_start[L], cnt, key, n, c->lev_c.p,
									 c->noff.p, c->n_first.p, c->n_end.p, c->n_parent.p,
									 c->p_father.p, c->s_pm.p, c->s_vh.p, c->s_type.p, c->s_lev.p, c->n_sm.p,
									 c->n_cl.p, c->n_aux.p, c->n_child.p, c->dom, dp);
	else
	  k_moments<<<gl_grid(cnt, 128), 128, 0, c->stream>>>(L, c->lev_list.p + c->lev_start[L], cnt, key, n, c->lev_c.p,
							     c->noff.p, c->n_first.p, c->n_end.p, c->n_parent.p,
							     c->p_father.p, c->s_pm.p, c->s_vh.p, c->s_type.p, c->s_lev.p, c->n_sm.p,
							     c->n_cl.p, c->n_aux.p, c->n_child.p, c->dom, dp);
	GL_LAUNCH_COUNT(c);
      }
  if(entries)
    {
      size_t ne = (size_t) c->nentries + 1;
      GL_CUDA(c, c->e_next.ensure(ne));
      GL_CUDA(c, c->e_A.ensure(ne));
      GL_CUDA(c, c->e_B.ensure(ne));
      GL_CUDA(c, c->e_C.ensure(ne));
      GL_CUDA(c, c->e_hmax.ensure(ne));
      GL_CUDA(c, c->e_rng.ensure(ne));
      GL_CUDA(c, c->e_F0.ensure(ne));
      GL_CUDA(c, c->e_F1.ensure(ne));
      GL_CUDA(c, c->e_F2.ensure(ne));
      GL_CUDA(c, c->grav_scale.ensure(2));
      k_grav_scale<<<1, 32, 0, c->stream>>>(m, n, c->n_sm.p, c->s_pm.p, c->dom.len, c->grav_scale.p);
      k_entries_particles<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->noff.p, c->s_pm.p, c->s_vh.p, c->s_type.p, c->s_lev.p, key,
								  c->e_next.p, c->e_A.p, c->e_B.p, c->e_C.p, c->e_hmax.p, c->e_rng.p,
								  c->e_F0.p, c->e_F1.p, c->e_F2.p, c->grav_scale.p, c->dom, dp);
      c->kernel_launches += 2;
      if(m > 0)
	{
	  k_entries_nodes<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, n, c->noff.p, c->n_first.p, c->n_end.p, c->n_sm.p,
								   c->n_cl.p, c->n_aux.p, c->e_next.p, c->e_A.p, c->e_B.p,
								   c->e_C.p, c->e_hmax.p, c->e_F0.p, c->e_F1.p, c->e_F2.p, c->grav_scale.p,
								   c->dom);
	  GL_LAUNCH_COUNT(c);
	  k_entries_range<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, c->n_first.p, c->n_end.p, c->n_aux.p, c->e_rng.p);
	  GL_LAUNCH_COUNT(c);
	}
      if(glk_apply_velpred(c)) return 1;	// after the moments: node vmax is built from P[].Vel (forcetree.c:684)
    }
  cudaEventRecord(c->ev[3], c->stream);
  GL_CUDA(c, cudaGetLastError());
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_keys_sort = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
  c->st.ms_treebuild = ms;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
  c->st.ms_moments = ms;
  c->st.numnodes = m;
  c->st.numentries = c->nentries;
  c->tree_valid = entries;
  c->tree_flag = !entries;
  c->tgt_dirty = true;
  return 0;
}

// ---------------------------------------------------------------- K9 hmax refresh ----
// forcetree.c:1661-1705: from every active gas particle climb the father chain while the particle's
// Hsml raises the node's hmax.  Positive doubles order like their bit patterns, so an unsigned
// 64-bit atomicMax implements "if(h > hmax) hmax = h" without a lock.
__global__ void k_update_hmax(long long n, const int *__restrict__ s_type, const double4 *__restrict__ s_vh,
			      const int *__restrict__ noff, const int *__restrict__ p_father, const int *__restrict__ n_parent,
			      const int *__restrict__ n_first, double *__restrict__ e_hmax)
{
  // every gas particle of the tree, not only this call's targets: particles whose Hsml did not change
  // stop at their father (old >= h), and after a multi-GPU allgather of Hsml the particles owned by
  // other GPUs must raise hmax here as well (the reference exchanges the raised top-leaf values,
  // forcetree.c:1722-1747)
  long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(p >= n) return;
  const int ty = s_type[p];
  if(ty != 0 && ty != -1) return;
  // a cut cell (type -1, multi-rank mode) climbs with the hmax its owner sent after density (its own entry)
  double h = (ty == 0) ? s_vh[p].w : e_hmax[p + noff[p + 1]];
  unsigned long long hb = (unsigned long long) __double_as_longlong(h);
  int no = p_father[p];
  while(no >= 0)
    {
      unsigned long long *addr = (unsigned long long *) &e_hmax[(long long) n_first[no] + no];
      unsigned long long old = atomicMax(addr, hb);
      if(old >= hb) break;
      no = n_parent[no];
    }
}

int glt_launch_update_hmax(gl_ctx *c)
{
  k_update_hmax<<<gl_grid(c->nit, 256), 256, 0, c->stream>>>(c->nit, c->s_type.p, c->s_vh.p, c->noff.p, c->p_father.p, c->n_parent.p,
							      c->n_first.p, c->e_hmax.p);
  GL_LAUNCH_COUNT(c);
  return 0;
}

__global__ void k_debug_hsml_consistency(long long nit, long long n, const int *__restrict__ perm, const int *__restrict__ s_type,
					 const double4 *__restrict__ s_vh, const double *__restrict__ in_hsml, int *__restrict__ bad)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nit) return;
  const int j = perm[i];
  if(j >= n || s_type[i] != 0) return;
  if(s_vh[i].w != in_hsml[j]) atomicAdd(bad, 1);
}

int glt_update_hmax(gl_ctx *c)
{
  if(!c->tree_valid) return c->fail("force_update_hmax: no tree");
  cudaEventRecord(c->ev[0], c->stream);
  if(c->comm && c->comm->world > 1)
    if(gll_exchange_sph_state(c)) return 1;	// multi-rank: Hsml + SphP of the imported particles, hmax of the cut cells
  if(glt_launch_update_hmax(c)) return 1;
  if(getenv("GL_DEBUG_LET"))
    {
      int *bad = c->scan_tmp_i.p + 63, hb = 0;
      cudaMemsetAsync(bad, 0, sizeof(int), c->stream);
      k_debug_hsml_consistency<<<gl_grid(c->nit, 256), 256, 0, c->stream>>>(c->nit, c->n, c->perm.p, c->s_type.p, c->s_vh.p, c->in_hsml.p, bad);
      cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
      cudaStreamSynchronize(c->stream);
      fprintf(stderr, "[gl] rank %d: force_update_hmax: %d gas particles whose tree-order Hsml differs from the input-order Hsml (n %lld, own %lld)\n",
	      c->comm ? c->comm->rank : 0, hb, c->n, c->n_own);
    }
  cudaEventRecord(c->ev[1], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_hmax = ms;
  return 0;
}

__global__ void k_node_table(long long m, const int *__restrict__ n_first, const signed char *__restrict__ n_level,
			     const double4 *__restrict__ n_sm, const double4 *__restrict__ n_cl,
			     const double4 *__restrict__ n_aux, const double *__restrict__ e_hmax, double *__restrict__ out)
{
  long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(id >= m) return;
  double4 sm = n_sm[id], cl = n_cl[id], ax = n_aux[id];
  double *o = out + 12 * id;
  o[0] = cl.w; o[1] = cl.x; o[2] = cl.y; o[3] = cl.z;
  o[4] = sm.w; o[5] = sm.x; o[6] = sm.y; o[7] = sm.z;
  o[8] = ax.x; o[9] = e_hmax[(long long) n_first[id] + id]; o[10] = ax.z; o[11] = (double) n_level[id];
}

int glt_get_nodes(gl_ctx *c, double *table, long long cap)
{
  if(!c->tree_valid) return c->fail("gl_get_nodes: no tree");
  long long m = c->nnodes;
  if(cap < m) return c->fail("gl_get_nodes: buffer too small");
  if(m == 0) return 0;
  DBuf<double> tmp;
  GL_CUDA(c, tmp.ensure((size_t) m * 12));
  k_node_table<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, c->n_first.p, c->n_level.p, c->n_sm.p, c->n_cl.p, c->n_aux.p,
							c->e_hmax.p, tmp.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaMemcpyAsync(table, tmp.p, (size_t) m * 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  tmp.release();
  return 0;
}

// ---------------------------------------------------------------- tree export (reference layout) ----
// reference index of DFS entry x: a particle is its P[] index, cell k is MaxPart + k, the end of the tree is -1
__device__ __forceinline__ int ref_index(long long x, long long nentries, const int *__restrict__ e_next,
					 const int2 *__restrict__ e_rng, const int *__restrict__ perm, int maxpart)
{
  if(x >= nentries) return -1;
  const int nx = e_next[x];
  if(nx < 0) return perm[(-1 - nx) >> 3];
  return maxpart + (int) (x - e_rng[x].x);
}

__global__ void k_export_nodes(long long m, long long nentries, gl_tree_layout L, int maxpart, int ti_current,
			       const int *__restrict__ n_first, const int *__restrict__ n_parent,
			       const double4 *__restrict__ n_sm, const double4 *__restrict__ n_cl,
			       const double4 *__restrict__ n_aux, const int *__restrict__ e_next,
			       const int2 *__restrict__ e_rng, const double *__restrict__ e_hmax,
			       const int *__restrict__ perm, char *__restrict__ nodes, char *__restrict__ ext)
{
  const long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(id >= m) return;
  const long long e = (long long) n_first[id] + id;
  char *nd = nodes + id * L.node_stride, *ex = ext + id * L.ext_stride;
  const double4 cl = n_cl[id], sm = n_sm[id], ax = n_aux[id];
  *(double *) (nd + L.n_len) = cl.w;
  double *ctr = (double *) (nd + L.n_center);
  ctr[0] = cl.x; ctr[1] = cl.y; ctr[2] = cl.z;
  if(L.n_maxsoft >= 0) *(double *) (nd + L.n_maxsoft) = ax.x;
  double *sp = (double *) (nd + L.n_s);
  sp[0] = sm.x; sp[1] = sm.y; sp[2] = sm.z;
  *(double *) (nd + L.n_mass) = sm.w;
  *(unsigned int *) (nd + L.n_bitflags) = L.bitflags_value;
  *(int *) (nd + L.n_nextnode) = ref_index(e + 1, nentries, e_next, e_rng, perm, maxpart);	// first daughter follows in DFS order
  *(int *) (nd + L.n_sibling) = ref_index(e_next[e], nentries, e_next, e_rng, perm, maxpart);
  *(int *) (nd + L.n_father) = (n_parent[id] >= 0) ? maxpart + n_parent[id] : -1;
  *(int *) (nd + L.n_Ti_current) = ti_current;
  // extNODE: vs / divVmax follow from k_export_velocity_moments; dp starts at zero (forcetree.c:742-744)
  *(double *) (ex + L.e_vmax) = ax.z;
  *(double *) (ex + L.e_hmax) = e_hmax[e];
  *(int *) (ex + L.e_Ti_lastkicked) = ti_current;
}

// extNODE.vs / .divVmax (forcetree.c:522-545 initial values, :571-573 / :609-610 child cells, :650-652 / :679-681
// particles, :704-706 division by the mass): the mass-weighted velocity of the cell in child-slot order, and the largest
// SphP.v.DivVel below it.  Only force_drift_node / force_kick_node read them (the dynamic tree update), so this pass runs
// in the export and not in the build: one launch per level, bottom-up, over the child table the moment pass left.
__global__ void k_export_velocity_moments(int count, const int *__restrict__ lev_list, const int *__restrict__ n_child,
					  const int *__restrict__ n_first, const int *__restrict__ n_end, const int2 *__restrict__ e_rng,
					  const double4 *__restrict__ n_sm, const double4 *__restrict__ s_pm, const int *__restrict__ s_type,
					  const int *__restrict__ perm, const double *__restrict__ in_vel, const double *__restrict__ divvel,
					  double4 *__restrict__ vsd)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= count) return;
  const int id = lev_list[t];
  const int *tab = n_child + 8ll * id;
  double v0 = 0, v1 = 0, v2 = 0, dmax = 0;
  auto particle = [&](int p) {
    const int i = perm[p];
    const double m = s_pm[p].w;
    v0 = __dadd_rn(v0, __dmul_rn(m, in_vel[3ll * i]));
    v1 = __dadd_rn(v1, __dmul_rn(m, in_vel[3ll * i + 1]));
    v2 = __dadd_rn(v2, __dmul_rn(m, in_vel[3ll * i + 2]));
    if(s_type[p] == 0 && divvel && divvel[i] > dmax) dmax = divvel[i];
  };
  if(tab[0] == INT_MIN)
    for(int p = n_first[id]; p < n_end[id]; p++) particle(p);	// finest-level cell with coincident keys: every particle hangs directly below
  else
    for(int k = 0; k < 8; k++)
      {
	const int code = tab[k];
	if(code == -1) break;
	if(code <= -2)
	  particle(-(code + 2));
	else
	  {
	    const int child = code - e_rng[code].x;	// DFS entry -> cell index
	    const double m = n_sm[child].w;
	    const double4 cv = vsd[child];
	    v0 = __dadd_rn(v0, __dmul_rn(m, cv.x));
	    v1 = __dadd_rn(v1, __dmul_rn(m, cv.y));
	    v2 = __dadd_rn(v2, __dmul_rn(m, cv.z));
	    if(cv.w > dmax) dmax = cv.w;
	  }
      }
  const double mass = n_sm[id].w;
  if(mass != 0)
    {
      v0 = __ddiv_rn(v0, mass); v1 = __ddiv_rn(v1, mass); v2 = __ddiv_rn(v2, mass);
    }
  else
    v0 = v1 = v2 = 0;
  vsd[id] = make_double4(v0, v1, v2, dmax);
}

__global__ void k_export_vs_store(long long m, gl_tree_layout L, const double4 *__restrict__ vsd, char *__restrict__ ext)
{
  const long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(id >= m) return;
  char *ex = ext + id * L.ext_stride;
  const double4 v = vsd[id];
  if(L.e_vs >= 0)
    {
      double *o = (double *) (ex + L.e_vs);
      o[0] = v.x; o[1] = v.y; o[2] = v.z;
    }
  if(L.e_divVmax >= 0) *(double *) (ex + L.e_divVmax) = v.w;
}

__global__ void k_export_particles(long long n, long long nentries, int maxpart, const int *__restrict__ noff,
				   const int *__restrict__ p_father, const int *__restrict__ e_next,
				   const int2 *__restrict__ e_rng, const int *__restrict__ perm, int *__restrict__ nextnode,
				   int *__restrict__ father)
{
  const long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(p >= n) return;
  const int i = perm[p];
  const long long e = p + noff[p + 1];
  nextnode[i] = ref_index(e + 1, nentries, e_next, e_rng, perm, maxpart);
  father[i] = (p_father[p] >= 0) ? maxpart + p_father[p] : -1;
}

extern "C" int gl_export_tree(gl_ctx *c, const gl_tree_layout *lay, void *nodes_base, void *extnodes_base, int *nextnode,
			      int *father, long long MaxPart, long long max_nodes, int Ti_Current, long long *numnodes)
{
  cudaSetDevice(c->device);
  if(!c->tree_valid) return c->fail("gl_export_tree: no tree (call gl_force_treebuild or gl_gravity_tree first)");
  if(c->ncell > 0 || (c->comm && c->comm->world > 1)) return c->fail("gl_export_tree: single rank only (the reference inserts pseudo-particles here, forcetree.c:439)");
  const long long n = c->n, m = c->nnodes;
  if(n > MaxPart) return c->fail("gl_export_tree: more particles than MaxPart");
  if(m > max_nodes) return c->fail("gl_export_tree: more cells than MaxNodes (the reference would retry with a larger TreeAllocFactor, forcetree.c:71-91)");
  if(lay->node_stride <= 0 || lay->ext_stride <= 0) return c->fail("gl_export_tree: bad layout");
  DBuf<char> dn, de;
  DBuf<int> dx;
  GL_CUDA(c, dn.ensure((size_t) std::max<long long>(m, 1) * lay->node_stride));
  GL_CUDA(c, de.ensure((size_t) std::max<long long>(m, 1) * lay->ext_stride));
  GL_CUDA(c, dx.ensure(2 * (size_t) n));
  GL_CUDA(c, cudaMemsetAsync(dn.p, 0, (size_t) m * lay->node_stride, c->stream));
  GL_CUDA(c, cudaMemsetAsync(de.p, 0, (size_t) m * lay->ext_stride, c->stream));
  if(m > 0)
    {
      k_export_nodes<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, c->nentries, *lay, (int) MaxPart, Ti_Current, c->n_first.p, c->n_parent.p,
							    c->n_sm.p, c->n_cl.p, c->n_aux.p, c->e_next.p, c->e_rng.p, c->e_hmax.p,
							    c->perm.p, dn.p, de.p);
      GL_LAUNCH_COUNT(c);
    }
  DBuf<double4> vsd;
  if(m > 0 && (lay->e_vs >= 0 || lay->e_divVmax >= 0))
    {
      GL_CUDA(c, vsd.ensure((size_t) m));
      const double *divv = c->o_divvel.cap >= (size_t) n ? c->o_divvel.p : nullptr;	// SphP.v.DivVel as the last density() left it
      for(int L = GL_BITS; L >= 0; L--)
	{
	  const int cnt = c->lev_start[L + 1] - c->lev_start[L];
	  if(cnt <= 0) continue;
	  k_export_velocity_moments<<<gl_grid(cnt, 128), 128, 0, c->stream>>>(cnt, c->lev_list.p + c->lev_start[L], c->n_child.p, c->n_first.p,
									      c->n_end.p, c->e_rng.p, c->n_sm.p, c->s_pm.p, c->s_type.p, c->perm.p,
									      c->in_vel.p, divv, vsd.p);
	  GL_LAUNCH_COUNT(c);
	}
      k_export_vs_store<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, *lay, vsd.p, de.p);
      GL_LAUNCH_COUNT(c);
    }
  k_export_particles<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->nentries, (int) MaxPart, c->noff.p, c->p_father.p, c->e_next.p,
							      c->e_rng.p, c->perm.p, dx.p, dx.p + n);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaMemcpyAsync(nodes_base, dn.p, (size_t) m * lay->node_stride, cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(extnodes_base, de.p, (size_t) m * lay->ext_stride, cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(nextnode, dx.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(father, dx.p + n, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  dn.release();
  de.release();
  dx.release();
  vsd.release();
  if(numnodes) *numnodes = m;
  return 0;
}
