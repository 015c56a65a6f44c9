// gl_let.cu - multi-rank force path: the locally-essential tree (LET) exchange.
//
// What the reference does.  Every MPI rank owns the particles of some Peano-Hilbert key ranges (domain.c:1075-1234),
// builds the tree of its own particles below a replicated "top tree", all-gathers the moments of the top-tree leaves
// (force_exchange_pseudodata forcetree.c:879-1016, re-summed in slot order by force_treeupdate_pseudos :1025) and
// hangs the foreign leaves in as pseudo-particles (force_insert_pseudo_particles :439).  A walk that must open a
// pseudo-particle EXPORTS THE TARGET to the owner and gets a partial result back (gravtree.c:175-347,
// density.c:193-342, hydra.c:274-520: two MPI_Sendrecv rounds per phase plus the count matrices).
//
// What happens here instead (one GPU per rank, NCCL inside the library).  Sources move, targets never do:
//   1. every rank builds the tree of its OWN particles (same keys: the cube is the all-reduced extent);
//   2. it all-gathers a description of its targets - the Peano-Hilbert key hull, and GL_LET_BOXES bounding boxes of
//      consecutive targets with the smallest aold = ErrTolForceAcc * OldAcc and the largest search radius in each;
//   3. for every other rank it walks its own tree top-down and EXPORTS a cell as a whole ("cut cell": moments only)
//      where (a) no other rank has a particle inside it - its key range misses their hulls, so the moments are the
//      global ones, bit for bit - and (b) no target box of that rank can open it (conservative forms of
//      forcetree.c:2076 / :2085 / :2094-2105 / :2124-2137) or reach it with a neighbour search (ngb.c:136-152,
//      :270-289, radius padded by let->fs); otherwise it descends; particles reached that way are exported as
//      particles.  One personalised all-to-all moves the cells, one the particles;
//   4. the receiver appends the imported particles behind its own, sorts the cut cells in as ITEMS of the Morton
//      order (gl_tree.cu) and builds the merged tree.  Every cell above the cut cells is re-summed from its children in
//      slot order by the ordinary moment pass - the force_treeupdate_pseudos step, exact for the same reason.
// The merged tree IS the global tree wherever a target of this rank can look, so interaction counts, neighbour
// sets and every sum are those of a single-rank run.  That is enforced, not assumed: a walk that wants to open a cut
// cell, or a search sphere that touches one, raises GL_ERR_LET_VIOLATION, all ranks agree on it (one 8-byte
// all-gather per phase) and repeat the phase with every particle exported (a full replication, always exact).
// After density one more exchange over the same lists carries the new Hsml and the SphP fields hydro reads from
// neighbours, and the new hmax of the cut cells (the reference's forcetree.c:1722-1747 exchange).
#include "gl_internal.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <climits>
#include <chrono>

gl_dparams gl_make_dparams(const gl_ctx *c);

#define GL_LET_BOXES 64		// target boxes per rank (plus one for the whole rank)
#define GL_LET_BOXDBL 14		// lo[3], hi[3], aold_min, hsph_max (-1: no SPH target inside), reach lo[3], hi[3]: bbox of the search spheres
#define GL_LET_HDR (2 * GL_LET_BOXES)	// per-rank header: the Peano-Hilbert key hull {min, max} of every stretch (bit patterns)
#define GL_LET_SUB 8		// every stretch is cut into GL_LET_SUB sub-stretches with boxes of their own (second level of the test)
#define GL_LET_LEAVES (GL_LET_BOXES * GL_LET_SUB)
// one rank's block: [hulls | box 0 = whole rank, boxes 1..GL_LET_BOXES = stretches | GL_LET_LEAVES sub-stretch boxes]
#define GL_LET_BLK (GL_LET_HDR + (GL_LET_BOXES + 1 + GL_LET_LEAVES) * GL_LET_BOXDBL)
#define GL_LET_LEAF0 (GL_LET_HDR + (GL_LET_BOXES + 1) * GL_LET_BOXDBL)
#define GL_LET_PART_DBL 10	// exported particle: pos[3], vel[3], mass, hsml, entropy, (type | timebin << 32)
#define GL_LET_CELL_DBL 8	// exported cell: key (bits), s[3], mass, maxsoft, hmax, level + 32 * vmax-less flag
#define GL_LET_SPH_DBL 6	// second exchange, per particle: hsml, density, pressure, dhsml, divvel, curlvel

struct gl_let_state
{
  bool full = false;		// export every particle (after a violation, for the rest of this particle load)
  bool built_bh = false, built_grav = false;
  double fs = 1.3;		// padding of the search radii in the export test: one 1.26 step of the Hsml iteration (density.c:610-642) stays inside
  DBuf<double> box_own, box_all;
  DBuf<unsigned long long> ranges;
  DBuf<unsigned char> complete, state, xflag;
  DBuf<int> sel, bounds;
  DBuf<double> send_c, send_p, recv_c, recv_p, send_s, recv_s, send_h, recv_h;
  DBuf<int> exp_part;		// tree index (own tree) -> kept as INPUT index of every exported particle, destination-major
  DBuf<unsigned long long> exp_ckey;
  DBuf<int> exp_clev;
  DBuf<unsigned long long> flagbuf;
  std::vector<size_t> sc_off, sc_cnt, sp_off, sp_cnt, rc_off, rc_cnt, rp_off, rp_cnt;	// in records, per rank
  long long n_exp_c = 0, n_exp_p = 0, n_imp_c = 0, n_imp_p = 0;
  gl_let_stats st{};
};

static gl_let_state *let_of(gl_ctx *c)
{
  if(!c->let) c->let = new gl_let_state();
  return c->let;
}

void gll_release(gl_ctx *c)
{
  if(!c->let) return;
  gl_let_state *L = c->let;
  L->box_own.release(); L->box_all.release(); L->ranges.release(); L->complete.release(); L->state.release();
  L->xflag.release(); L->sel.release(); L->bounds.release(); L->send_c.release(); L->send_p.release(); L->recv_c.release();
  L->recv_p.release(); L->send_s.release(); L->recv_s.release(); L->send_h.release(); L->recv_h.release();
  L->exp_part.release(); L->exp_ckey.release(); L->exp_clev.release(); L->flagbuf.release();
  delete L;
  c->let = nullptr;
}

#define GL_COMM(c, call)                                                              \
  do {                                                                                \
    if((call)) return (c)->fail(std::string("multi-rank exchange: ") + (c)->comm->err); \
  } while(0)

// ---------------------------------------------------------------- target boxes ----
// block b > 0: the b-th of GL_LET_BOXES stretches of the own particles in the caller's order (the reference keeps
// P[] in Peano-Hilbert order, peano.c:25, so a stretch is a compact clump); block 0: all of them
__global__ void k_let_boxes(long long n_own, const long long *__restrict__ bounds, const double *__restrict__ pos,
			    const double *__restrict__ oldacc, const double *__restrict__ hsml, const int *__restrict__ type,
			    double errtolforceacc, double fs, double *__restrict__ boxes)
{
  // block = (stretch b, sub-stretch q): a box per sub-stretch; stretch boxes and the whole-rank box are unions (k_let_box_union)
  const int b = blockIdx.x / GL_LET_SUB, q = blockIdx.x % GL_LET_SUB;
  const long long s0 = bounds[b], s1 = bounds[b + 1];
  const long long i0 = s0 + (s1 - s0) * q / GL_LET_SUB, i1 = s0 + (s1 - s0) * (q + 1) / GL_LET_SUB;
  (void) n_own;
  // v[0..2] min pos, [3..5] max pos, [6] min aold, [7] max search radius, [8..10] min (pos - fs h), [11..13] max (pos + fs h)
  double v[GL_LET_BOXDBL];
  for(int k = 0; k < 3; k++)
    {
      v[k] = v[8 + k] = 1e300;
      v[3 + k] = v[11 + k] = -1e300;
    }
  v[6] = 1e300;
  v[7] = -1.0;
  for(long long i = i0 + threadIdx.x; i < i1; i += blockDim.x)
    {
      const int ty = type[i];
      const bool sphtarget = ty == 0 || ty == 5 || ty == 2;	// density_isactive types (density.c:1035)
      const double h = sphtarget ? hsml[i] : 0.0;
      for(int k = 0; k < 3; k++)
	{
	  const double x = pos[3 * i + k];
	  v[k] = fmin(v[k], x);
	  v[3 + k] = fmax(v[3 + k], x);
	  if(sphtarget)
	    {
	      v[8 + k] = fmin(v[8 + k], x - fs * h);
	      v[11 + k] = fmax(v[11 + k], x + fs * h);
	    }
	}
      v[6] = fmin(v[6], __dmul_rn(errtolforceacc, oldacc[i]));
      if(sphtarget) v[7] = fmax(v[7], h);
    }
  __shared__ double sh[8][GL_LET_BOXDBL];
  for(int o = 16; o > 0; o >>= 1)
    for(int k = 0; k < GL_LET_BOXDBL; k++)
      {
	const double w = __shfl_xor_sync(0xffffffffu, v[k], o);
	const bool ismin = k < 3 || k == 6 || (k >= 8 && k < 11);
	v[k] = ismin ? fmin(v[k], w) : fmax(v[k], w);
      }
  const int w = threadIdx.x >> 5;
  if((threadIdx.x & 31) == 0)
    for(int k = 0; k < GL_LET_BOXDBL; k++) sh[w][k] = v[k];
  __syncthreads();
  if(threadIdx.x == 0)
    {
      for(int ww = 1; ww < blockDim.x / 32; ww++)
	for(int k = 0; k < GL_LET_BOXDBL; k++)
	  {
	    const bool ismin = k < 3 || k == 6 || (k >= 8 && k < 11);
	    sh[0][k] = ismin ? fmin(sh[0][k], sh[ww][k]) : fmax(sh[0][k], sh[ww][k]);
	  }
      double *o = boxes + GL_LET_LEAF0 + (long long) blockIdx.x * GL_LET_BOXDBL;
      for(int k = 0; k < GL_LET_BOXDBL; k++) o[k] = sh[0][k];
    }
}

// block b >= 1: stretch b = union of its sub-stretch boxes; then (second launch, block 0) the whole rank = union of the stretches
__global__ void k_let_box_union(double *__restrict__ boxes, int whole)
{
  const int k = threadIdx.x;
  if(k >= GL_LET_BOXDBL) return;
  const bool ismin = k < 3 || k == 6 || (k >= 8 && k < 11);
  double v = ismin ? 1e300 : -1e300;
  if(k == 7) v = -1.0;
  if(whole)
    {
      for(int b = 1; b <= GL_LET_BOXES; b++)
	{
	  const double *o = boxes + GL_LET_HDR + (long long) b * GL_LET_BOXDBL;
	  if(o[0] > o[3]) continue;	// empty stretch
	  v = ismin ? fmin(v, o[k]) : fmax(v, o[k]);
	}
      boxes[GL_LET_HDR + k] = v;
      return;
    }
  const int b = blockIdx.x;	// stretch b -> box b + 1
  for(int q = 0; q < GL_LET_SUB; q++)
    {
      const double *o = boxes + GL_LET_LEAF0 + (long long) (b * GL_LET_SUB + q) * GL_LET_BOXDBL;
      if(o[0] > o[3]) continue;
      v = ismin ? fmin(v, o[k]) : fmax(v, o[k]);
    }
  boxes[GL_LET_HDR + (long long) (b + 1) * GL_LET_BOXDBL + k] = v;
}

__global__ void k_let_count_gas(long long n, const int *__restrict__ type, int *__restrict__ count)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  const bool g = i < n && type[i] == 0;
  const unsigned b = __ballot_sync(0xffffffffu, g);
  if((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// ---------------------------------------------------------------- export selection ----
struct LetSel
{
  int world, me, bh, want_grav, full, want_sph;
  double theta, fs;
};

// Can a target inside `box` open this cell, or reach it with a neighbour search?  Conservative: a "no" is certain.
__device__ __forceinline__ bool box_may_need(const double *__restrict__ box, const double4 &sm, const double4 &cl, const double4 &ax,
					     const LetSel &P)
{
  if(box[0] > box[3]) return false;	// empty stretch
  if(P.want_grav)
    {
      const double d0 = fmax(fmax(box[0] - sm.x, sm.x - box[3]), 0.0), d1 = fmax(fmax(box[1] - sm.y, sm.y - box[4]), 0.0),
	d2z = fmax(fmax(box[2] - sm.z, sm.z - box[5]), 0.0);
      const double d2 = d0 * d0 + d1 * d1 + d2z * d2z;
      const double len2 = cl.w * cl.w;
      if(P.bh)
	{
	  if(len2 >= d2 * P.theta * P.theta * (1 - 1e-6)) return true;	// forcetree.c:2076
	}
      else
	{
	  if(sm.w * len2 >= d2 * d2 * box[6] * (1 - 1e-6)) return true;	// forcetree.c:2085 with the smallest aold of the box
	  // forcetree.c:2094-2105: some target position may lie in the 0.6 len cube
	  const double q = 0.6001 * cl.w;
	  const bool inside = !(box[0] >= cl.x + q || box[3] <= cl.x - q || box[1] >= cl.y + q || box[4] <= cl.y - q ||
				box[2] >= cl.z + q || box[5] <= cl.z - q);
	  if(inside) return true;
	}
      if(ax.x > 0 && d2 <= ax.x * ax.x * (1 + 1e-6)) return true;	// forcetree.c:2124-2137 softening guard
    }
  if(P.want_sph && ax.y > 0 && box[7] >= 0)
    {
      // gas below the cell and SPH targets in the box
      const double half = 0.5 * cl.w * 1.000001;
      // (1) some target's search sphere r < h_i (ngb.c:270-289; hydra.c:1266 first half) may reach the cell: the cell
      // meets the bounding box of the spheres (radii padded by fs, formed per target in k_let_boxes)
      const bool meets = !(box[8] > cl.x + half || box[11] < cl.x - half || box[9] > cl.y + half || box[12] < cl.y - half ||
			   box[10] > cl.z + half || box[13] < cl.z - half);
      if(meets) return true;
      // (2) some particle of the cell may reach a target with ITS radius, r < h_j <= hmax (ngb.c:136-152)
      const double R = P.fs * ax.y;
      const double e0 = fmax(fmax(box[0] - (cl.x + half), (cl.x - half) - box[3]), 0.0),
	e1 = fmax(fmax(box[1] - (cl.y + half), (cl.y - half) - box[4]), 0.0),
	e2z = fmax(fmax(box[2] - (cl.z + half), (cl.z - half) - box[5]), 0.0);
      if(e0 * e0 + e1 * e1 + e2z * e2z <= R * R) return true;
    }
  return false;
}

// one level of the top-down pass: blockIdx.y = destination rank (its target boxes are staged in shared memory),
// EIGHT LANES per cell of that level: lane q looks at the stretches q+1, q+9, ... (and, below a stretch that may
// need the cell, at its sub-stretches); the eight leave together as soon as one of them found a box that needs the cell
// (the work per cell varies from one test to several hundred - one thread per cell left most of a warp idle).
// state: 0 = descend (children are looked at), 1 = exported as a cut cell, 2 = below an exported cell
__global__ void __launch_bounds__(128)
k_let_select_level(int count, const int *__restrict__ lev_list, long long m, const int *__restrict__ n_parent,
		   const double4 *__restrict__ n_sm, const double4 *__restrict__ n_cl, const double4 *__restrict__ n_aux,
		   const unsigned char *__restrict__ complete, const double *__restrict__ box_all,
		   unsigned char *__restrict__ state, LetSel P)
{
  __shared__ double sbox[(GL_LET_BOXES + 1) * GL_LET_BOXDBL];
  const int d = blockIdx.y;
  if(d == P.me) return;
  const double *rb = box_all + (long long) d * GL_LET_BLK + GL_LET_HDR;
  const double *leaves = box_all + (long long) d * GL_LET_BLK + GL_LET_LEAF0;
  for(int k = threadIdx.x; k < (GL_LET_BOXES + 1) * GL_LET_BOXDBL; k += blockDim.x) sbox[k] = rb[k];
  __syncthreads();
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, q = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);	// my group of eight lanes
  const bool live = t < count;
  const int id = live ? lev_list[t] : 0;
  unsigned char *st = state + (long long) d * m;
  const int par = live ? n_parent[id] : -1;
  int res = 0;			// all eight lanes of a cell compute the same res
  bool test = false;
  if(live)
    {
      if(par >= 0 && st[par] != 0)
	res = 2;
      else if(!P.full && complete[id])
	test = true;
    }
  if(test)
    {
      const double4 sm = n_sm[id], cl = n_cl[id], ax = n_aux[id];
      bool need = box_may_need(sbox, sm, cl, ax, P);	// the whole rank first: far ranks are settled by one test
      if(need)
	{
	  need = false;
	  for(int it = 0; it < GL_LET_BOXES / 8; it++)
	    {
	      const int b = it * 8 + q + 1;
	      bool mine = false;
	      if(box_may_need(sbox + b * GL_LET_BOXDBL, sm, cl, ax, P))
		for(int k = 0; k < GL_LET_SUB && !mine; k++)
		  mine = box_may_need(leaves + (long long) ((b - 1) * GL_LET_SUB + k) * GL_LET_BOXDBL, sm, cl, ax, P);
	      if(__ballot_sync(gmask, mine) & gmask)
		{
		  need = true;
		  break;
		}
	    }
	}
      res = need ? 0 : 1;
    }
  if(live && q == 0) st[id] = (unsigned char) res;
}

// selection flags: per destination M cell flags followed by n particle flags
__global__ void k_let_flags(long long m, long long n, int world, int me, const unsigned char *__restrict__ state,
			    const int *__restrict__ p_father, unsigned char *__restrict__ xflag)
{
  const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  const long long per = m + n;
  if(t >= per * world) return;
  const int d = (int) (t / per);
  const long long k = t % per;
  unsigned char f = 0;
  if(d != me)
    {
      const unsigned char *st = state + (long long) d * m;
      if(k < m)
	f = st[k] == 1;
      else
	{
	  const int fa = p_father[k - m];
	  f = (fa < 0) ? 1 : (st[fa] == 0);	// a particle directly below a cell that is descended into
	}
    }
  xflag[t] = f;
}

struct FlagSet
{
  const unsigned char *f;
  __device__ bool operator()(const int &i) const { return f[i] != 0; }
};

// first position in sel[0..ns) with value >= v
__global__ void k_let_bounds(const int *__restrict__ sel, int ns, int world, long long m, long long per, int *__restrict__ bounds)
{
  const int t = threadIdx.x;
  if(t > 2 * world) return;
  const long long v = (t == 2 * world) ? per * world : (long long) (t / 2) * per + ((t & 1) ? m : 0);
  int lo = 0, hi = ns;
  while(lo < hi)
    {
      const int mid = lo + (hi - lo) / 2;
      if(sel[mid] < v)
	lo = mid + 1;
      else
	hi = mid;
    }
  bounds[t] = lo;
}

__global__ void k_let_pack(int ns, const int *__restrict__ sel, const int *__restrict__ bounds, int world, long long m, long long per,
			   const int *__restrict__ n_first, const signed char *__restrict__ n_level, const unsigned long long *__restrict__ key,
			   const double4 *__restrict__ n_sm, const double4 *__restrict__ n_aux, const int *__restrict__ perm,
			   const double *__restrict__ pos, const double *__restrict__ vel, const double *__restrict__ mass,
			   const double *__restrict__ hsml, const double *__restrict__ entropy, const int *__restrict__ type,
			   const int *__restrict__ timebin, const double *__restrict__ velpred, long long n_velpred,
			   double *__restrict__ send_c, double *__restrict__ send_p,
			   unsigned long long *__restrict__ exp_ckey, int *__restrict__ exp_clev, int *__restrict__ exp_part)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= ns) return;
  const long long g = sel[k];
  const int d = (int) (g / per);
  const long long idx = g % per;
  // cells of all destinations are packed first (in destination order), then the particles
  if(idx < m)
    {
      // position among the cells: cells of destinations < d, plus rank inside this destination's cell block
      int before = 0;
      for(int q = 0; q < d; q++) before += bounds[2 * q + 1] - bounds[2 * q];
      const int pc = before + (k - bounds[2 * d]);
      const int L = n_level[idx];
      const int sh = 3 * (GL_BITS - L);
      const unsigned long long ka = key[n_first[idx]];
      const unsigned long long low = (sh >= 63) ? 0ull : ((ka >> sh) << sh);
      const double4 sm = n_sm[idx], ax = n_aux[idx];
      double *o = send_c + (long long) pc * GL_LET_CELL_DBL;
      o[0] = __longlong_as_double((long long) low);
      o[1] = sm.x; o[2] = sm.y; o[3] = sm.z; o[4] = sm.w;
      o[5] = ax.x; o[6] = ax.y;
      o[7] = (double) L;
      exp_ckey[pc] = low;
      exp_clev[pc] = L;
    }
  else
    {
      int before = 0;
      for(int q = 0; q < d; q++) before += bounds[2 * q + 2] - bounds[2 * q + 1];
      const int pp = before + (k - bounds[2 * d + 1]);
      const int j = perm[idx - m];	// input index of the own particle
      double *o = send_p + (long long) pp * GL_LET_PART_DBL;
      o[0] = pos[3ll * j]; o[1] = pos[3ll * j + 1]; o[2] = pos[3ll * j + 2];
      // what a neighbour reads is SphP.VelPred for gas (density.c:772, hydra.c:962); the drop-in layer keeps it apart from P.Vel
      const double *vv = (velpred && j < n_velpred && type[j] == 0) ? velpred : vel;
      o[3] = vv[3ll * j]; o[4] = vv[3ll * j + 1]; o[5] = vv[3ll * j + 2];
      o[6] = mass[j];
      o[7] = hsml[j];
      o[8] = entropy[j];
      o[9] = __longlong_as_double(((long long) (unsigned) type[j]) | (((long long) timebin[j]) << 32));
      exp_part[pp] = j;
    }
}

__global__ void k_let_unpack_particles(long long ni, long long base, const double *__restrict__ rec, double *__restrict__ pos,
				       double *__restrict__ vel, double *__restrict__ mass, double *__restrict__ hsml,
				       double *__restrict__ oldacc, double *__restrict__ entropy, double *__restrict__ dtentropy,
				       int *__restrict__ type, int *__restrict__ timebin, int *__restrict__ tibeg)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= ni) return;
  const double *r = rec + i * GL_LET_PART_DBL;
  const long long j = base + i;
  pos[3 * j] = r[0]; pos[3 * j + 1] = r[1]; pos[3 * j + 2] = r[2];
  vel[3 * j] = r[3]; vel[3 * j + 1] = r[4]; vel[3 * j + 2] = r[5];
  mass[j] = r[6];
  hsml[j] = r[7];
  entropy[j] = r[8];
  const long long tt = __double_as_longlong(r[9]);
  type[j] = (int) (tt & 0xffffffffll);
  timebin[j] = (int) (tt >> 32);
  oldacc[j] = 0;
  dtentropy[j] = 0;
  tibeg[j] = 0;
}

__global__ void k_let_unpack_cells(long long nc, const double *__restrict__ rec, unsigned long long *__restrict__ c_key,
				   signed char *__restrict__ c_level, double4 *__restrict__ c_sm, double4 *__restrict__ c_aux)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nc) return;
  const double *r = rec + i * GL_LET_CELL_DBL;
  c_key[i] = (unsigned long long) __double_as_longlong(r[0]);
  c_sm[i] = make_double4(r[1], r[2], r[3], r[4]);
  c_aux[i] = make_double4(r[5], r[6], 0.0, 0.0);	// maxsoft, hmax (vmax only feeds the dynamic tree update: not sent)
  c_level[i] = (signed char) (int) r[7];
}

__global__ void k_iota(int *p, long long n, int base);

static int exchange_counts(gl_ctx *c, gl_let_state *L, const std::vector<long long> &mine, std::vector<long long> &all)
{
  // every rank's row of (cells, particles) per destination -> the full matrix on every rank
  const int W = c->comm->world;
  GL_CUDA(c, L->flagbuf.ensure((size_t) 2 * W * (W + 1) + 16));
  unsigned long long *own = L->flagbuf.p, *full = L->flagbuf.p + 2 * W;
  GL_CUDA(c, cudaMemcpyAsync(own, mine.data(), sizeof(long long) * 2 * W, cudaMemcpyHostToDevice, c->stream));
  GL_COMM(c, c->comm->allgather(own, full, sizeof(long long) * 2 * W, c->stream));
  all.resize((size_t) 2 * W * W);
  GL_CUDA(c, cudaMemcpyAsync(all.data(), full, sizeof(long long) * 2 * W * W, cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// All ranks agree on "somebody saw a violation" (one 8-byte all-gather).  Returns 0/1 in *any.
int gll_agree_violation(gl_ctx *c, int mine, int *any)
{
  gl_let_state *L = let_of(c);
  const int W = c->comm->world;
  GL_CUDA(c, L->flagbuf.ensure((size_t) 2 * W * (W + 1) + 16));
  unsigned long long v = (unsigned long long) mine, *own = L->flagbuf.p, *full = L->flagbuf.p + 2 * W;
  GL_CUDA(c, cudaMemcpyAsync(own, &v, sizeof(v), cudaMemcpyHostToDevice, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_COMM(c, c->comm->allgather(own, full, sizeof(v), c->stream));
  std::vector<unsigned long long> h(W);
  GL_CUDA(c, cudaMemcpyAsync(h.data(), full, sizeof(v) * W, cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  *any = 0;
  for(int r = 0; r < W; r++)
    if(h[r]) *any = 1;
  return 0;
}

void gll_set_full(gl_ctx *c, bool full) { let_of(c)->full = full; }
void gll_set_padding(gl_ctx *c, double fs) { let_of(c)->fs = fs; }
// was the current merged tree selected with the criterion the coming walk uses?
bool gll_built_for(gl_ctx *c, bool bh, bool grav)
{
  gl_let_state *L = let_of(c);
  return c->tree_valid && (!grav || (L->built_grav && L->built_bh == bh));
}

int gll_distributed_build(gl_ctx *c)
{
  gl_let_state *L = let_of(c);
  gl_comm *cm = c->comm;
  const int W = cm->world, me = cm->rank;
  if(W > GL_LET_MAXRANKS) return c->fail("multi-rank build: more than 64 ranks");
  if(c->n_own <= 0) return c->fail("multi-rank build: no own particles loaded (gl_load_particles)");
  cudaEvent_t e0 = c->ev[4], e1 = c->ev[5], e2 = c->ev[6], e3 = c->ev[7];
  cudaEventRecord(e0, c->stream);
  // GL_DEBUG_LET_TIME: wall-clock of every sub-step (with a device synchronisation in front of each stamp)
  static const bool dbgt = getenv("GL_DEBUG_LET_TIME") != nullptr;
  auto tnow = [&]() {
    cudaStreamSynchronize(c->stream);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  double tl = dbgt ? tnow() : 0;
  auto stamp = [&](const char *what) {
    if(!dbgt) return;
    const double t = tnow();
    if(me == 0) fprintf(stderr, "[gl let] %-28s %8.3f ms\n", what, t - tl);
    tl = t;
  };
  // ---- 1. own particles only, global cube (domain.c:1996-1997: MPI_Allreduce of the extent) ----
  c->n = c->n_own;
  c->ncell = 0;
  {
    GL_CUDA(c, c->dom_dev.ensure((size_t) 148 * 4 * 6 + 16 + 6 * (size_t) W));
    if(glk_local_minmax(c, c->dom_dev.p)) return 1;
    double *all = c->dom_dev.p + 148 * 4 * 6 + 16;
    GL_COMM(c, cm->allgather(c->dom_dev.p, all, 6 * sizeof(double), c->stream));
    std::vector<double> h(6 * (size_t) W);
    GL_CUDA(c, cudaMemcpyAsync(h.data(), all, sizeof(double) * 6 * W, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(c, cudaStreamSynchronize(c->stream));
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for(int r = 0; r < W; r++)
      for(int k = 0; k < 3; k++)
	{
	  lo[k] = std::min(lo[k], h[6 * r + k]);
	  hi[k] = std::max(hi[k], h[6 * r + 3 + k]);
	}
    if(!c->domain_imposed) glk_set_domain_minmax(c, lo, hi);
  }
  stamp("extent all-reduce");
  // ---- 2. the tree of the own particles (no walk records) ----
  const bool keep_ordered = c->tgt_ordered;
  const long long keep_tgt = c->n_tgt_in;
  c->tgt_ordered = true;	// no Peano-Hilbert sort of the own particles: they are the caller's ordered segment
  if(c->n_tgt_in < 0) c->n_tgt_in = 0;
  if(glt_build_local(c, false)) return 1;
  c->tgt_ordered = keep_ordered;
  c->n_tgt_in = keep_tgt;
  const long long n = c->n_own, m = c->nnodes, per = m + n;
  cudaEventRecord(e1, c->stream);
  stamp("own tree");
  // ---- 3. target description of every rank ----
  const size_t blk = GL_LET_BLK;
  GL_CUDA(c, L->box_own.ensure(blk));
  GL_CUDA(c, L->box_all.ensure(blk * W));
  GL_CUDA(c, L->ranges.ensure(2 * (size_t) W + 2));
  GL_CUDA(c, cudaMemsetAsync(L->box_own.p, 0, blk * sizeof(double), c->stream));
  {
    // stretches of the caller's order: the gas block (P[0..N_gas), allvars.h) and the rest apart, so that a sink far
    // from the rank's gas gets a box and a key hull of its own instead of inflating a gas stretch
    GL_CUDA(c, L->bounds.ensure(2 * (size_t) W + 4 + 2 * (GL_LET_BOXES + 2)));
    long long *bdev = (long long *) (L->bounds.p + 2 * W + 4);
    int *gcount = L->bounds.p + 2 * W + 3;
    GL_CUDA(c, cudaMemsetAsync(gcount, 0, sizeof(int), c->stream));
    k_let_count_gas<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->in_type.p, gcount);
    GL_LAUNCH_COUNT(c);
    int ngas = 0;
    GL_CUDA(c, cudaMemcpyAsync(&ngas, gcount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(c, cudaStreamSynchronize(c->stream));
    const long long nrest = n - ngas;
    int nb_rest = 0;
    if(nrest > 0 && ngas > 0) nb_rest = (int) std::min<long long>(GL_LET_BOXES - 8, std::max<long long>(8, (GL_LET_BOXES * nrest) / n));
    if(ngas == 0) nb_rest = GL_LET_BOXES;
    const int nb_gas = GL_LET_BOXES - nb_rest;
    long long hbnd[GL_LET_BOXES + 1];
    for(int b = 0; b <= nb_gas; b++) hbnd[b] = nb_gas ? (long long) b * ngas / nb_gas : 0;
    for(int b = 1; b <= nb_rest; b++) hbnd[nb_gas + b] = ngas + (long long) b * nrest / nb_rest;
    hbnd[GL_LET_BOXES] = n;
    GL_CUDA(c, cudaMemcpyAsync(bdev, hbnd, sizeof(hbnd), cudaMemcpyHostToDevice, c->stream));
    // key hulls: per stretch of the caller's order (tight while that order is the Peano-Hilbert order of this step's grid),
    // or from the sorted keys when the rank has kept its particles over several steps (gl_run_steps)
    if(c->let_gap_hulls ? glk_ph_gap_hulls(c, GL_LET_BOXES, (unsigned long long *) L->box_own.p)
       : glk_ph_chunk_hulls(c, bdev, GL_LET_BOXES, (unsigned long long *) L->box_own.p))
      return 1;
    k_let_boxes<<<GL_LET_LEAVES, 128, 0, c->stream>>>(n, bdev, c->in_pos.p, c->in_oldacc.p, c->in_hsml.p, c->in_type.p,
						       c->par.ErrTolForceAcc, L->fs, L->box_own.p);
    k_let_box_union<<<GL_LET_BOXES, 32, 0, c->stream>>>(L->box_own.p, 0);
    k_let_box_union<<<1, 32, 0, c->stream>>>(L->box_own.p, 1);
    c->kernel_launches += 3;
    GL_CUDA(c, cudaStreamSynchronize(c->stream));	// hbnd is a stack array
  }
  GL_COMM(c, cm->allgather(L->box_own.p, L->box_all.p, blk * sizeof(double), c->stream));
  stamp("boxes + hulls + all-gather");
  // ---- 4. what each other rank needs from here ----
  GL_CUDA(c, L->complete.ensure((size_t) m + 1));
  GL_CUDA(c, L->state.ensure((size_t) W * (m + 1)));
  GL_CUDA(c, L->xflag.ensure((size_t) W * per + 1));
  if(glk_node_complete(c, (const unsigned long long *) L->box_all.p, W, GL_LET_BOXES, (long long) blk, me, L->complete.p)) return 1;
  LetSel P;
  P.world = W; P.me = me; P.bh = c->par.ErrTolTheta != 0; P.want_grav = c->let_want_grav ? 1 : 0; P.full = L->full ? 1 : 0;
  if(getenv("GL_DEBUG_LET_FULL")) P.full = 1;
  P.theta = c->par.ErrTolTheta; P.fs = L->fs;
  P.want_sph = getenv("GL_DEBUG_LET_NOSPH") ? 0 : 1;	// (debug switches: results are then wrong or violations follow)
  if(getenv("GL_DEBUG_LET_NOGRAV")) P.want_grav = 0;
  for(int lev = 0; lev <= GL_BITS && m > 0; lev++)
    {
      const int cnt = c->lev_start[lev + 1] - c->lev_start[lev];
      if(cnt <= 0) continue;
      k_let_select_level<<<dim3(gl_grid(8ll * cnt, 128), W), 128, 0, c->stream>>>(cnt, c->lev_list.p + c->lev_start[lev], m, c->n_parent.p,
									    c->n_sm.p, c->n_cl.p, c->n_aux.p, L->complete.p,
									    L->box_all.p, L->state.p, P);
      GL_LAUNCH_COUNT(c);
    }
  stamp("complete + level passes");
  k_let_flags<<<gl_grid(per * W, 256), 256, 0, c->stream>>>(m, n, W, me, L->state.p, c->p_father.p, L->xflag.p);
  GL_LAUNCH_COUNT(c);
  if(per * W >= (1ll << 31)) return c->fail("multi-rank build: selection index space exceeds 2^31");
  GL_CUDA(c, L->sel.ensure((size_t) per * W + 1));
  int ns = 0;
  {
    size_t tmp = 0;
    int *dcount = L->bounds.p + 2 * W + 2;
    cub::CountingInputIterator<int> it(0);
    FlagSet fs{L->xflag.p};
    cub::DeviceSelect::If(nullptr, tmp, it, L->sel.p, dcount, (int) (per * W), fs, c->stream);
    GL_CUDA(c, c->cub_tmp.ensure(tmp));
    GL_CUDA(c, cub::DeviceSelect::If(c->cub_tmp.p, tmp, it, L->sel.p, dcount, (int) (per * W), fs, c->stream));
    c->kernel_launches += 2;
    GL_CUDA(c, cudaMemcpyAsync(&ns, dcount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  std::vector<int> hb(2 * (size_t) W + 1, 0);
  k_let_bounds<<<1, 2 * W + 1 <= 32 ? 32 : 2 * W + 1, 0, c->stream>>>(L->sel.p, ns, W, m, per, L->bounds.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaMemcpyAsync(hb.data(), L->bounds.p, sizeof(int) * (2 * W + 1), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<long long> mine(2 * (size_t) W), all;
  long long tc = 0, tp = 0;
  L->sc_off.assign(W, 0); L->sc_cnt.assign(W, 0); L->sp_off.assign(W, 0); L->sp_cnt.assign(W, 0);
  for(int d = 0; d < W; d++)
    {
      const long long nc = hb[2 * d + 1] - hb[2 * d], np = hb[2 * d + 2] - hb[2 * d + 1];
      mine[2 * d] = nc;
      mine[2 * d + 1] = np;
      L->sc_off[d] = tc; L->sc_cnt[d] = nc;
      L->sp_off[d] = tp; L->sp_cnt[d] = np;
      tc += nc;
      tp += np;
    }
  L->n_exp_c = tc;
  L->n_exp_p = tp;
  GL_CUDA(c, L->send_c.ensure((size_t) tc * GL_LET_CELL_DBL + 8));
  GL_CUDA(c, L->send_p.ensure((size_t) tp * GL_LET_PART_DBL + 8));
  GL_CUDA(c, L->exp_ckey.ensure((size_t) tc + 1));
  GL_CUDA(c, L->exp_clev.ensure((size_t) tc + 1));
  GL_CUDA(c, L->exp_part.ensure((size_t) tp + 1));
  if(ns > 0)
    {
      k_let_pack<<<gl_grid(ns, 256), 256, 0, c->stream>>>(ns, L->sel.p, L->bounds.p, W, m, per, c->n_first.p, c->n_level.p, c->key_m_sorted.p,
							   c->n_sm.p, c->n_aux.p, c->perm.p, c->in_pos.p, c->in_vel.p, c->in_mass.p, c->in_hsml.p,
							   c->in_entropy.p, c->in_type.p, c->in_timebin.p, c->have_velpred ? c->in_velpred.p : nullptr,
							   c->have_velpred ? c->n_velpred : 0, L->send_c.p, L->send_p.p,
							   L->exp_ckey.p, L->exp_clev.p, L->exp_part.p);
      GL_LAUNCH_COUNT(c);
    }
  cudaEventRecord(e2, c->stream);
  stamp("flags + compaction + pack");
  // ---- 5. the exchange ----
  if(exchange_counts(c, L, mine, all)) return 1;
  L->rc_off.assign(W, 0); L->rc_cnt.assign(W, 0); L->rp_off.assign(W, 0); L->rp_cnt.assign(W, 0);
  long long ic = 0, ip = 0;
  for(int s = 0; s < W; s++)
    {
      const long long nc = all[(size_t) 2 * W * s + 2 * me], np = all[(size_t) 2 * W * s + 2 * me + 1];
      L->rc_off[s] = ic; L->rc_cnt[s] = nc;
      L->rp_off[s] = ip; L->rp_cnt[s] = np;
      ic += nc;
      ip += np;
    }
  L->n_imp_c = ic;
  L->n_imp_p = ip;
  if(n + ip >= (1ll << 27) + (1ll << 24)) return c->fail("multi-rank build: own + imported particles exceed the 2^27 design limit");
  GL_CUDA(c, L->recv_c.ensure((size_t) ic * GL_LET_CELL_DBL + 8));
  GL_CUDA(c, L->recv_p.ensure((size_t) ip * GL_LET_PART_DBL + 8));
  {
    std::vector<size_t> so(W), sc(W), ro(W), rc(W);
    for(int r = 0; r < W; r++)
      {
	so[r] = L->sc_off[r] * GL_LET_CELL_DBL * sizeof(double); sc[r] = L->sc_cnt[r] * GL_LET_CELL_DBL * sizeof(double);
	ro[r] = L->rc_off[r] * GL_LET_CELL_DBL * sizeof(double); rc[r] = L->rc_cnt[r] * GL_LET_CELL_DBL * sizeof(double);
      }
    GL_COMM(c, cm->alltoallv(L->send_c.p, so.data(), sc.data(), L->recv_c.p, ro.data(), rc.data(), c->stream));
    for(int r = 0; r < W; r++)
      {
	so[r] = L->sp_off[r] * GL_LET_PART_DBL * sizeof(double); sc[r] = L->sp_cnt[r] * GL_LET_PART_DBL * sizeof(double);
	ro[r] = L->rp_off[r] * GL_LET_PART_DBL * sizeof(double); rc[r] = L->rp_cnt[r] * GL_LET_PART_DBL * sizeof(double);
      }
    GL_COMM(c, cm->alltoallv(L->send_p.p, so.data(), sc.data(), L->recv_p.p, ro.data(), rc.data(), c->stream));
  }
  stamp("counts + two all-to-alls");
  // ---- 6. imported particles behind the own ones, cut cells as extra items ----
  const size_t nn = (size_t) (n + ip);
  GL_CUDA(c, c->in_pos.ensure_keep(3 * nn)); GL_CUDA(c, c->in_vel.ensure_keep(3 * nn)); GL_CUDA(c, c->in_mass.ensure_keep(nn));
  GL_CUDA(c, c->in_hsml.ensure_keep(nn)); GL_CUDA(c, c->in_oldacc.ensure_keep(nn)); GL_CUDA(c, c->in_entropy.ensure_keep(nn));
  GL_CUDA(c, c->in_dtentropy.ensure_keep(nn)); GL_CUDA(c, c->in_type.ensure_keep(nn)); GL_CUDA(c, c->in_timebin.ensure_keep(nn));
  GL_CUDA(c, c->in_tibegstep.ensure_keep(nn));
  if(ip > 0)
    {
      k_let_unpack_particles<<<gl_grid(ip, 256), 256, 0, c->stream>>>(ip, n, L->recv_p.p, c->in_pos.p, c->in_vel.p, c->in_mass.p, c->in_hsml.p,
								       c->in_oldacc.p, c->in_entropy.p, c->in_dtentropy.p, c->in_type.p,
								       c->in_timebin.p, c->in_tibegstep.p);
      GL_LAUNCH_COUNT(c);
    }
  GL_CUDA(c, c->c_key.ensure((size_t) ic + 1)); GL_CUDA(c, c->c_level.ensure((size_t) ic + 1));
  GL_CUDA(c, c->c_sm.ensure((size_t) ic + 1)); GL_CUDA(c, c->c_aux.ensure((size_t) ic + 1));
  if(ic > 0)
    {
      k_let_unpack_cells<<<gl_grid(ic, 256), 256, 0, c->stream>>>(ic, L->recv_c.p, c->c_key.p, c->c_level.p, c->c_sm.p, c->c_aux.p);
      GL_LAUNCH_COUNT(c);
    }
  c->n = n + ip;
  c->ncell = ic;
  // the targets of this rank are its own particles, in the caller's order
  if(c->n_tgt_in < 0 || c->let_auto_targets)
    {
      GL_CUDA(c, c->tgt_in.ensure((size_t) n + 1));
      k_iota<<<gl_grid(n, 256), 256, 0, c->stream>>>(c->tgt_in.p, n, 0);
      GL_LAUNCH_COUNT(c);
      c->n_tgt_in = n;
      c->tgt_ordered = true;
      c->let_auto_targets = true;
      c->tgt_host.clear();
    }
  c->res_first = 0;
  c->res_count = n;
  cudaEventRecord(e3, c->stream);
  stamp("unpack imports");
  // ---- 7. the merged tree ----
  if(c->n_tgt_in >= 0) c->tgt_ordered = true;	// a caller's list is taken as it comes (results do not depend on the order)
  if(glt_build_local(c, true)) return 1;
  stamp("merged tree");
  L->built_bh = P.bh != 0;
  L->built_grav = P.want_grav != 0;
  float t01, t12, t23;
  cudaEventElapsedTime(&t01, e0, e1);
  cudaEventElapsedTime(&t12, e1, e2);
  cudaEventElapsedTime(&t23, e2, e3);
  L->st.ms_own_tree = t01;
  L->st.ms_select = t12;
  L->st.ms_exchange = t23;
  L->st.exported_cells = tc; L->st.exported_particles = tp;
  L->st.imported_cells = ic; L->st.imported_particles = ip;
  L->st.own_particles = n;
  L->st.full = L->full ? 1 : 0;
  L->st.builds++;
  return 0;
}

// ---------------------------------------------------------------- after density: SphP state of the imports ----
__global__ void k_let_pack_sph(long long np, const int *__restrict__ exp_part, const double *__restrict__ hsml,
			       const double *__restrict__ density, const double *__restrict__ pressure, const double *__restrict__ dhsml,
			       const double *__restrict__ divvel, const double *__restrict__ curlvel, double *__restrict__ out)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= np) return;
  const int j = exp_part[i];
  double *o = out + i * GL_LET_SPH_DBL;
  o[0] = hsml[j]; o[1] = density[j]; o[2] = pressure[j]; o[3] = dhsml[j]; o[4] = divvel[j]; o[5] = curlvel[j];
}

// position of the first sorted item whose key is >= k
__device__ __forceinline__ long long item_lower_bound(const unsigned long long *__restrict__ key, long long nit, unsigned long long k)
{
  long long lo = 0, hi = nit;
  while(lo < hi)
    {
      const long long mid = lo + (hi - lo) / 2;
      if(key[mid] < k)
	lo = mid + 1;
      else
	hi = mid;
    }
  return lo;
}

// hmax of an exported cell in the CURRENT (merged) tree of its owner: the cell is the node (a, L), a = first item
__global__ void k_let_pack_hmax(long long nc, const unsigned long long *__restrict__ ckey, const int *__restrict__ clev,
				const unsigned long long *__restrict__ key, long long nit, const signed char *__restrict__ lev_c,
				const int *__restrict__ noff, const double *__restrict__ e_hmax, double *__restrict__ out,
				unsigned long long *__restrict__ counter)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nc) return;
  const long long a = item_lower_bound(key, nit, ckey[i]);
  const int L = clev[i];
  double h = 0;
  if(a < nit && (int) lev_c[a] < L)
    {
      const long long id = noff[a] + (L - (int) lev_c[a] - 1);
      if(id < noff[a + 1])
	h = e_hmax[a + id];
      else
	atomicMax(&counter[2], 7002ull);
    }
  else
    atomicMax(&counter[2], 7002ull);
  out[i] = h;
}

__global__ void k_let_unpack_sph(long long ni, long long base, const double *__restrict__ rec, const int *__restrict__ inv,
				 double4 *__restrict__ s_vh, double *__restrict__ hsml, double *__restrict__ density,
				 double *__restrict__ pressure, double *__restrict__ dhsml, double *__restrict__ divvel,
				 double *__restrict__ curlvel)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= ni) return;
  const double *r = rec + i * GL_LET_SPH_DBL;
  const long long j = base + i;
  hsml[j] = r[0];
  s_vh[inv[j]].w = r[0];
  density[j] = r[1]; pressure[j] = r[2]; dhsml[j] = r[3]; divvel[j] = r[4]; curlvel[j] = r[5];
}

__global__ void k_let_unpack_hmax(long long nc, const double *__restrict__ rec, const unsigned long long *__restrict__ c_key,
				  const unsigned long long *__restrict__ key, long long nit, const int *__restrict__ noff,
				  double4 *__restrict__ s_vh, double *__restrict__ e_hmax)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nc) return;
  const long long a = item_lower_bound(key, nit, c_key[i]);	// the cut cell's own item (nothing else lies inside the cell)
  if(a >= nit) return;
  s_vh[a].y = rec[i];
  e_hmax[a + noff[a + 1]] = rec[i];
}

int gll_exchange_sph_state(gl_ctx *c)
{
  gl_let_state *L = let_of(c);
  gl_comm *cm = c->comm;
  const int W = cm->world;
  if(!c->tree_valid) return c->fail("force_update_hmax: no tree");
  if(c->o_density.cap < (size_t) c->n_own) return c->fail("force_update_hmax (multi-rank): density() has not been run");
  const long long n = c->n_own;
  GL_CUDA(c, c->o_density.ensure_keep((size_t) c->n)); GL_CUDA(c, c->o_pressure.ensure_keep((size_t) c->n));
  GL_CUDA(c, c->o_dhsml.ensure_keep((size_t) c->n)); GL_CUDA(c, c->o_divvel.ensure_keep((size_t) c->n));
  GL_CUDA(c, c->o_curlvel.ensure_keep((size_t) c->n));
  c->let_need_sph_exchange = false;
  cudaEventRecord(c->ev[4], c->stream);
  // own new Hsml into the cells first: the hmax of an exported cell is read off this tree
  if(glt_launch_update_hmax(c)) return 1;
  GL_CUDA(c, L->send_s.ensure((size_t) L->n_exp_p * GL_LET_SPH_DBL + 8));
  GL_CUDA(c, L->recv_s.ensure((size_t) L->n_imp_p * GL_LET_SPH_DBL + 8));
  GL_CUDA(c, L->send_h.ensure((size_t) L->n_exp_c + 8));
  GL_CUDA(c, L->recv_h.ensure((size_t) L->n_imp_c + 8));
  GL_CUDA(c, c->d_counter.ensure(16));
  if(L->n_exp_p > 0)
    {
      k_let_pack_sph<<<gl_grid(L->n_exp_p, 256), 256, 0, c->stream>>>(L->n_exp_p, L->exp_part.p, c->in_hsml.p, c->o_density.p, c->o_pressure.p,
								       c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p, L->send_s.p);
      GL_LAUNCH_COUNT(c);
    }
  if(L->n_exp_c > 0)
    {
      k_let_pack_hmax<<<gl_grid(L->n_exp_c, 256), 256, 0, c->stream>>>(L->n_exp_c, L->exp_ckey.p, L->exp_clev.p, c->key_m_sorted.p, c->nit,
									c->lev_c.p, c->noff.p, c->e_hmax.p, L->send_h.p, c->d_counter.p);
      GL_LAUNCH_COUNT(c);
    }
  std::vector<size_t> so(W), sc(W), ro(W), rc(W);
  for(int r = 0; r < W; r++)
    {
      so[r] = L->sp_off[r] * GL_LET_SPH_DBL * sizeof(double); sc[r] = L->sp_cnt[r] * GL_LET_SPH_DBL * sizeof(double);
      ro[r] = L->rp_off[r] * GL_LET_SPH_DBL * sizeof(double); rc[r] = L->rp_cnt[r] * GL_LET_SPH_DBL * sizeof(double);
    }
  GL_COMM(c, cm->alltoallv(L->send_s.p, so.data(), sc.data(), L->recv_s.p, ro.data(), rc.data(), c->stream));
  for(int r = 0; r < W; r++)
    {
      so[r] = L->sc_off[r] * sizeof(double); sc[r] = L->sc_cnt[r] * sizeof(double);
      ro[r] = L->rc_off[r] * sizeof(double); rc[r] = L->rc_cnt[r] * sizeof(double);
    }
  GL_COMM(c, cm->alltoallv(L->send_h.p, so.data(), sc.data(), L->recv_h.p, ro.data(), rc.data(), c->stream));
  if(L->n_imp_p > 0)
    {
      k_let_unpack_sph<<<gl_grid(L->n_imp_p, 256), 256, 0, c->stream>>>(L->n_imp_p, n, L->recv_s.p, c->inv.p, c->s_vh.p, c->in_hsml.p,
									 c->o_density.p, c->o_pressure.p, c->o_dhsml.p, c->o_divvel.p,
									 c->o_curlvel.p);
      GL_LAUNCH_COUNT(c);
    }
  if(L->n_imp_c > 0)
    {
      k_let_unpack_hmax<<<gl_grid(L->n_imp_c, 256), 256, 0, c->stream>>>(L->n_imp_c, L->recv_h.p, c->c_key.p, c->key_m_sorted.p, c->nit,
									  c->noff.p, c->s_vh.p, c->e_hmax.p);
      GL_LAUNCH_COUNT(c);
    }
  cudaEventRecord(c->ev[5], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
  L->st.ms_sph_exchange = ms;
  return 0;
}

extern "C" int gl_get_let_stats(gl_ctx *c, gl_let_stats *out)
{
  if(!c->let) return c->fail("gl_get_let_stats: no multi-rank build yet");
  *out = c->let->st;
  out->violations = c->let_violations;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Collective partition (SURVEY.md 8(e) "Partition"; domain.c:1075 domain_findSplit_work_balanced with ONE stretch per rank):
// the global Peano-Hilbert order is cut into `world` stretches of equal cost and every rank learns where each of its
// particles belongs.  No rank ever sees all keys: each sorts its own, describes its cost distribution by DS_SAMPLES keys at
// equal steps of its cumulative cost, the samples are all-gathered and every rank derives the same world-1 splitter keys
// from them (error per splitter <= total cost / DS_SAMPLES, 0.4 % of a stretch at 8 ranks).  The struct migration itself
// (domain_exchange, domain.c:665) stays with the caller, who owns the structs.
#define DS_SAMPLES 2048
namespace
{
__global__ void k_ds_iota(long long n, int *idx)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) idx[i] = (int) i;
}
__global__ void k_ds_gather_cost(long long n, const int *__restrict__ idx, const double *__restrict__ cost, double *__restrict__ out)
{
  const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k < n) out[k] = cost ? cost[idx[k]] : 1.0;
}
// sample s = the key at which the rank's cumulative cost reaches (s + 1/2) / DS_SAMPLES of its total
__global__ void k_ds_samples(long long n, const unsigned long long *__restrict__ key_sorted, const double *__restrict__ cum, unsigned long long *__restrict__ out)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if(s >= DS_SAMPLES) return;
  const double want = cum[n - 1] * ((s + 0.5) / DS_SAMPLES);
  long long lo = 0, hi = n - 1;
  while(lo < hi)
    {
      const long long mid = (lo + hi) >> 1;
      if(cum[mid] < want)
	lo = mid + 1;
      else
	hi = mid;
    }
  out[s] = key_sorted[lo];
}
__global__ void k_ds_dest(long long n, const unsigned long long *__restrict__ key, const unsigned long long *__restrict__ split, int nsplit,
			  int *__restrict__ dest, unsigned long long *__restrict__ count)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const unsigned long long k = key[i];
  int lo = 0, hi = nsplit;	// number of splitters <= k
  while(lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if(split[mid] <= k)
	lo = mid + 1;
      else
	hi = mid;
    }
  dest[i] = lo;
  atomicAdd(&count[lo], 1ull);
}
}	// namespace

extern "C" int gl_domain_split_collective(gl_ctx *c, const double *cost, int cost_where, int *dest, int dest_where, long long *to_go,
					   unsigned long long *splitters)
{
  cudaSetDevice(c->device);
  gl_comm *cm = c->comm;
  if(!cm) return c->fail("gl_domain_split_collective: no communicator (gl_comm_init)");
  const int W = cm->world, me = cm->rank;
  if(W > GL_LET_MAXRANKS) return c->fail("gl_domain_split_collective: more than 64 ranks");
  const long long n = c->n_own > 0 ? c->n_own : c->n;
  if(n <= 0) return c->fail("gl_domain_split_collective: no particles loaded");
  if(!dest || !to_go) return c->fail("gl_domain_split_collective: dest and to_go are required");
  c->n = n;			// (imports of an earlier build are not particles of this rank)
  c->ncell = 0;
  c->nit = n;
  // ---- the global cube: MPI_Allreduce of the extent (domain.c:1996-1997) ----
  {
    GL_CUDA(c, c->dom_dev.ensure((size_t) 148 * 4 * 6 + 16 + 6 * (size_t) W));
    if(glk_local_minmax(c, c->dom_dev.p)) return 1;
    double *all = c->dom_dev.p + 148 * 4 * 6 + 16;
    GL_COMM(c, cm->allgather(c->dom_dev.p, all, 6 * sizeof(double), c->stream));
    std::vector<double> h(6 * (size_t) W);
    GL_CUDA(c, cudaMemcpyAsync(h.data(), all, sizeof(double) * 6 * W, cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(c, cudaStreamSynchronize(c->stream));
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for(int r = 0; r < W; r++)
      for(int k = 0; k < 3; k++)
	{
	  lo[k] = std::min(lo[k], h[6 * r + k]);
	  hi[k] = std::max(hi[k], h[6 * r + 3 + k]);
	}
    if(!c->domain_imposed) glk_set_domain_minmax(c, lo, hi);
  }
  // ---- Peano-Hilbert keys of the own particles, sorted with the cost riding along ----
  if(glk_export_keys(c, nullptr, nullptr, 1)) return 1;	// -> c->key_ph (P[] order)
  DBuf<unsigned long long> ksort, samples, split_dev, count;
  DBuf<int> idx, idx_sorted, dest_dev;
  DBuf<double> cost_dev, cs, cum;
  struct Release
  {
    std::function<void()> f;
    ~Release() { f(); }
  } rel{ [&] { ksort.release(); samples.release(); split_dev.release(); count.release(); idx.release(); idx_sorted.release(); dest_dev.release();
	       cost_dev.release(); cs.release(); cum.release(); } };
  GL_CUDA(c, ksort.ensure((size_t) n)); GL_CUDA(c, idx.ensure((size_t) n)); GL_CUDA(c, idx_sorted.ensure((size_t) n));
  GL_CUDA(c, cs.ensure((size_t) n)); GL_CUDA(c, cum.ensure((size_t) n));
  GL_CUDA(c, samples.ensure((size_t) DS_SAMPLES * (W + 1))); GL_CUDA(c, split_dev.ensure((size_t) W)); GL_CUDA(c, count.ensure((size_t) W + 1));
  const double *cost_p = nullptr;
  if(cost)
    {
      if(cost_where)
	cost_p = cost;
      else
	{
	  GL_CUDA(c, cost_dev.ensure((size_t) n));
	  GL_CUDA(c, cudaMemcpyAsync(cost_dev.p, cost, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	  cost_p = cost_dev.p;
	}
    }
  k_ds_iota<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, idx.p);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->key_ph.p, ksort.p, idx.p, idx_sorted.p, (int) n, 0, 63, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->key_ph.p, ksort.p, idx.p, idx_sorted.p, (int) n, 0, 63, c->stream));
  k_ds_gather_cost<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, idx_sorted.p, cost_p, cs.p);
  tmp = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp, cs.p, cum.p, (int) n, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceScan::InclusiveSum(c->cub_tmp.p, tmp, cs.p, cum.p, (int) n, c->stream));
  // ---- samples + the rank's total, all-gathered ----
  unsigned long long *mine = samples.p + (size_t) DS_SAMPLES * W;
  k_ds_samples<<<gl_grid(DS_SAMPLES, 256), 256, 0, c->stream>>>(n, ksort.p, cum.p, mine);
  c->kernel_launches += 6;
  GL_COMM(c, cm->allgather(mine, samples.p, DS_SAMPLES * sizeof(unsigned long long), c->stream));
  double *tot_all = c->dom_dev.p;	// reuse: W doubles
  GL_COMM(c, cm->allgather(cum.p + (n - 1), tot_all, sizeof(double), c->stream));
  std::vector<unsigned long long> hs((size_t) DS_SAMPLES * W);
  std::vector<double> ht((size_t) W);
  GL_CUDA(c, cudaMemcpyAsync(hs.data(), samples.p, hs.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(ht.data(), tot_all, (size_t) W * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  // ---- the same splitters on every rank: weighted samples in key order, cut at k / world of the total ----
  std::vector<std::pair<unsigned long long, double> > ws;
  ws.reserve(hs.size());
  double total = 0;
  for(int r = 0; r < W; r++)
    {
      total += ht[(size_t) r];
      for(int s = 0; s < DS_SAMPLES; s++) ws.push_back({ hs[(size_t) r * DS_SAMPLES + s], ht[(size_t) r] / DS_SAMPLES });
    }
  std::stable_sort(ws.begin(), ws.end(), [](const std::pair<unsigned long long, double> &a, const std::pair<unsigned long long, double> &b) { return a.first < b.first; });
  std::vector<unsigned long long> split((size_t) W > 1 ? (size_t) W - 1 : 0);
  {
    double run = 0;
    size_t k = 0;
    for(int j = 1; j < W; j++)
      {
	const double want = total * j / W;
	while(k < ws.size() && run + ws[k].second <= want)
	  {
	    run += ws[k].second;
	    k++;
	  }
	// particles with a key >= the splitter go to rank j or beyond
	split[(size_t) j - 1] = k < ws.size() ? ws[k].first : ~0ull;
	if(j > 1 && split[(size_t) j - 1] < split[(size_t) j - 2]) split[(size_t) j - 1] = split[(size_t) j - 2];
      }
  }
  if(splitters)
    for(int j = 0; j + 1 < W; j++) splitters[j] = split[(size_t) j];
  // ---- destination of every own particle (P[] order) and the to-go counts ----
  GL_CUDA(c, dest_dev.ensure((size_t) n));
  if(W > 1) GL_CUDA(c, cudaMemcpyAsync(split_dev.p, split.data(), (size_t) (W - 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
  GL_CUDA(c, cudaMemsetAsync(count.p, 0, ((size_t) W + 1) * sizeof(unsigned long long), c->stream));
  k_ds_dest<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->key_ph.p, split_dev.p, W - 1, dest_dev.p, count.p);
  GL_LAUNCH_COUNT(c);
  std::vector<unsigned long long> hc((size_t) W);
  GL_CUDA(c, cudaMemcpyAsync(hc.data(), count.p, (size_t) W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(dest, dest_dev.p, (size_t) n * sizeof(int), dest_where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  for(int r = 0; r < W; r++) to_go[r] = r == me ? 0 : (long long) hc[(size_t) r];
  c->tree_valid = false;
  c->tree_flag = true;
  return 0;
}
