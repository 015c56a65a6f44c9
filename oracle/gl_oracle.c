/* gl_oracle.c - CPU restatement ("port") of the reference's force path, single rank, FP64.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build or execute this file; the product (libgl_b200.so and the
 * gadget_leicester_b200 package) never does, and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference ITSELF: oracle/refbuild compiles the unmodified
 * reference sources into oracle/_ref/ref_harness, tests/golden/make_golden.py ran it in the build
 * container and committed its outputs, and tests/test_oracle_golden.py requires this file to
 * reproduce them (interaction / pair counts exactly, FP64 fields to round-off).
 *
 * Every function names the reference lines it restates.  Compile: gcc -O2 -ffp-contract=off
 * (baseline x86-64, no FMA) - the same FP64 rounding as the reference build of SURVEY.md 8(c).
 *
 * usage: gl_oracle <in.glin> <out.glou> "<ops>"   (same container formats and op letters as
 *        oracle/refbuild/ref_harness.c: T S D R G U H N C K P A)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#define BITS 21			/* BITS_PER_DIMENSION allvars.h:58 */
#define MAXITER 150		/* allvars.h:139 */
#define GAMMA (7. / 5.)		/* allvars.h:64 */
#define GAMMA_MINUS1 (GAMMA - 1)
#define KC1 2.546479089470	/* allvars.h:247-253 */
#define KC2 15.278874536822
#define KC3 45.836623610466
#define KC4 30.557749073644
#define KC5 5.092958178941
#define KC6 (-15.278874536822)
#define NORM 4.188790204786
typedef unsigned long long peanokey;

/* ---- parameters (All.*) ---- */
#define NPAR 32
static double par[NPAR];
enum { P_THETA, P_ERRTOLFORCEACC, P_OPENCRIT, P_DESNUMNGB, P_MAXNGBDEV, P_ARTVISC, P_G, P_SOFT0,
  P_MINGASHSMLFRAC = P_SOFT0 + 6, P_TIMEBASE_INTERVAL, P_TI_CURRENT, P_BHNGBFACTOR, P_PARTALLOC
};
static double ErrTolTheta, ErrTolForceAcc, MaxNumNgbDeviation, ArtBulkViscConst, G, ForceSoftening[6], SofteningTable[6],
  MinGasHsml, Timebase_interval, BlackHoleNgbFactor;
static int TypeOfOpeningCriterion, DesNumNgb, Ti_Current;

/* ---- particles (P[] / SphP[] fields the path touches) ---- */
static int N, Ngas;
static double *Pos, *Vel, *Mass, *Hsml, *OldAcc, *Entropy, *DtEntropyIn;
static int *Type, *ID, *TimeBin;
static double *Potential;
static double *GravAccel, *GravCost, *NumNgb, *Density, *DhsmlDensityFactor, *DivVel, *CurlVel, *Pressure, *SinkDensity,
  *SinkEntropy, *SinkGasVel, *HydroAccel, *DtEntropy, *MaxSignalVel;

/* ---- tree (struct NODE / extNODE, allvars.h:1847,1887; Nextnode/Father :1915) ---- */
struct node
{
  double len, center[3], maxsoft;
  int suns[8];
  double s[3], mass, hmax, vmax;
  int multiple, sibling, nextnode, father;
};
static struct node *Nodes;	/* Nodes[no - N] for no >= N */
static int *Nextnode, *Father, MaxNodes, Numnodes, TreeReconstructFlag = 1, last;
static peanokey *MortonKey;
static double DomainLen, DomainFac, DomainCorner[3], DomainCenter[3];

static double now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------------------------------------------------------- keys ---- */
/* peano.c:300-316 restated through the curve's definition (visit rank of each octant and the frame
 * change into it) instead of the two 48x8 tables; identical keys (tests/test_oracle_golden.py). */
static const unsigned char rank_of[8] = { 0, 7, 1, 6, 3, 4, 2, 5 };
static const unsigned char child_col[8][3] = { {2, 1, 4}, {2, 1, 4}, {4, 2, 1}, {4, 2, 1}, {1, 2, 4}, {1, 2, 4}, {4, 2, 1}, {4, 2, 1} };
static const unsigned char child_mask[8] = { 0, 3, 0, 5, 6, 6, 0, 5 };
static unsigned aff(const unsigned char col[3], unsigned mask, unsigned p)
{
  unsigned v = mask, b;
  for(b = 0; b < 3; b++)
    if((p >> b) & 1) v ^= col[b];
  return v;
}
peanokey peano_hilbert_key(int x, int y, int z, int bits)
{
  unsigned char col[3] = { 1, 2, 4 }, ncol[3];
  unsigned mask = 0, nmask, b;
  peanokey key = 0;
  int i;
  for(i = bits - 1; i >= 0; i--)
    {
      unsigned pix = (((x >> i) & 1) << 2) | (((y >> i) & 1) << 1) | ((z >> i) & 1);
      unsigned q = aff(col, mask, pix);
      key = (key << 3) | rank_of[q];
      for(b = 0; b < 3; b++) ncol[b] = (unsigned char) (aff(child_col[q], child_mask[q], col[b]) ^ child_mask[q]);
      nmask = aff(child_col[q], child_mask[q], mask);
      memcpy(col, ncol, 3);
      mask = nmask;
    }
  return key;
}
peanokey morton_key(int x, int y, int z, int bits)	/* peano.c:320-332 */
{
  peanokey m = 0;
  int i;
  for(i = bits - 1; i >= 0; i--)
    m = (m << 3) + ((((z >> i) & 1) << 2) + (((y >> i) & 1) << 1) + ((x >> i) & 1));
  return m;
}

static void domain_findExtent(void)	/* domain.c:1972-2014 */
{
  double xmin[3] = { 1e300, 1e300, 1e300 }, xmax[3] = { -1e300, -1e300, -1e300 }, len = 0;
  int i, j;
  for(i = 0; i < N; i++)
    for(j = 0; j < 3; j++)
      {
	if(xmin[j] > Pos[3 * i + j]) xmin[j] = Pos[3 * i + j];
	if(xmax[j] < Pos[3 * i + j]) xmax[j] = Pos[3 * i + j];
      }
  for(j = 0; j < 3; j++)
    if(xmax[j] - xmin[j] > len) len = xmax[j] - xmin[j];
  len *= 1.001;
  for(j = 0; j < 3; j++)
    {
      DomainCenter[j] = 0.5 * (xmin[j] + xmax[j]);
      DomainCorner[j] = 0.5 * (xmin[j] + xmax[j]) - 0.5 * len;
    }
  DomainLen = len;
  DomainFac = 1.0 / len * (((peanokey) 1) << BITS);
}

/* ---------------------------------------------------------------- tree build ---- */
static void new_node(int no, int parent, int subnode)	/* forcetree.c:253-286 */
{
  struct node *nf = &Nodes[no - N], *pa = &Nodes[parent - N];
  double lenhalf = 0.25 * pa->len;
  int j;
  nf->len = 0.5 * pa->len;
  nf->center[0] = (subnode & 1) ? pa->center[0] + lenhalf : pa->center[0] - lenhalf;
  nf->center[1] = (subnode & 2) ? pa->center[1] + lenhalf : pa->center[1] - lenhalf;
  nf->center[2] = (subnode & 4) ? pa->center[2] + lenhalf : pa->center[2] - lenhalf;
  for(j = 0; j < 8; j++) nf->suns[j] = -1;
}

/* force_update_node_recursive, forcetree.c:468-872 (fields of the active build only) */
static void force_update_node_recursive(int no, int sib, int father)
{
  int j, jj, p, pp, nextsib, suns[8], count_particles = 0;
  double hmax = 0, vmax = 0, v, s[3] = { 0, 0, 0 }, mass = 0, maxsoft = 0;
  if(no >= N)
    {
      struct node *nd = &Nodes[no - N];
      for(j = 0; j < 8; j++) suns[j] = nd->suns[j];
      if(last >= 0)
	{
	  if(last >= N) Nodes[last - N].nextnode = no;
	  else Nextnode[last] = no;
	}
      last = no;
      for(j = 0; j < 8; j++)
	if((p = suns[j]) >= 0)
	  {
	    for(jj = j + 1; jj < 8; jj++)
	      if((pp = suns[jj]) >= 0) break;
	    nextsib = (jj < 8) ? pp : sib;
	    force_update_node_recursive(p, nextsib, no);
	    if(p >= N)
	      {
		struct node *c = &Nodes[p - N];
		mass += c->mass;
		s[0] += c->mass * c->s[0];
		s[1] += c->mass * c->s[1];
		s[2] += c->mass * c->s[2];
		if(c->mass > 0) count_particles += c->multiple ? 2 : 1;
		if(c->hmax > hmax) hmax = c->hmax;
		if(c->vmax > vmax) vmax = c->vmax;
		if(c->maxsoft > maxsoft) maxsoft = c->maxsoft;
	      }
	    else
	      {
		int k;
		count_particles++;
		mass += Mass[p];
		s[0] += Mass[p] * Pos[3 * p];
		s[1] += Mass[p] * Pos[3 * p + 1];
		s[2] += Mass[p] * Pos[3 * p + 2];
		if(Type[p] == 0)
		  {
		    if(Hsml[p] > hmax) hmax = Hsml[p];
		    if(Hsml[p] > maxsoft) maxsoft = Hsml[p];	/* forcetree.c:705-717 */
		  }
		else if(ForceSoftening[Type[p]] > maxsoft)
		  maxsoft = ForceSoftening[Type[p]];
		for(k = 0; k < 3; k++)
		  if((v = fabs(Vel[3 * p + k])) > vmax) vmax = v;
	      }
	  }
      if(mass)
	{
	  s[0] /= mass; s[1] /= mass; s[2] /= mass;
	}
      else
	{
	  s[0] = nd->center[0]; s[1] = nd->center[1]; s[2] = nd->center[2];
	}
      nd->mass = mass; nd->s[0] = s[0]; nd->s[1] = s[1]; nd->s[2] = s[2];
      nd->hmax = hmax; nd->vmax = vmax; nd->maxsoft = maxsoft;
      nd->multiple = count_particles > 1;
      nd->sibling = sib;
      nd->father = father;
    }
  else
    {
      if(last >= 0)
	{
	  if(last >= N) Nodes[last - N].nextnode = no;
	  else Nextnode[last] = no;
	}
      last = no;
      Father[no] = father;
    }
}

/* force_treebuild_single, forcetree.c:125-373, without the domain top-tree (single rank: it only adds
 * empty / single-particle cells that never interact, forcetree.c:1996-2004) and without the
 * sub-grid randomisation of :219-232 (inputs must not be degenerate; depth <= BITS is asserted). */
static int force_treebuild(void)
{
  int i, j, subnode = 0, shift, parent = -1, nfree, th, nn;
  peanokey morton;
  nfree = N;
  Nodes[0].len = DomainLen;
  for(j = 0; j < 3; j++) Nodes[0].center[j] = DomainCenter[j];
  for(j = 0; j < 8; j++) Nodes[0].suns[j] = -1;
  Numnodes = 1;
  nfree++;
  for(i = 0; i < N; i++)
    {
      morton = morton_key((int) ((Pos[3 * i] - DomainCorner[0]) * DomainFac), (int) ((Pos[3 * i + 1] - DomainCorner[1]) * DomainFac),
			  (int) ((Pos[3 * i + 2] - DomainCorner[2]) * DomainFac), BITS);
      MortonKey[i] = morton;
      shift = 3 * (BITS - 1);
      th = N;
      while(1)
	{
	  if(th >= N)
	    {
	      if(shift < 0) { fprintf(stderr, "gl_oracle: tree deeper than %d levels (coincident particles) - unsupported\n", BITS); exit(3); }
	      subnode = (int) ((morton >> shift) & 7);
	      nn = Nodes[th - N].suns[subnode];
	      shift -= 3;
	      if(nn >= 0)
		{
		  parent = th;
		  th = nn;
		}
	      else
		{
		  Nodes[th - N].suns[subnode] = i;
		  break;
		}
	    }
	  else
	    {
	      if(Numnodes >= MaxNodes) { fprintf(stderr, "gl_oracle: out of nodes\n"); exit(3); }
	      Nodes[parent - N].suns[subnode] = nfree;
	      new_node(nfree, parent, subnode);
	      if(shift < 0) { fprintf(stderr, "gl_oracle: tree deeper than %d levels - unsupported\n", BITS); exit(3); }
	      subnode = (int) ((MortonKey[th] >> shift) & 7);
	      Nodes[nfree - N].suns[subnode] = th;
	      th = nfree;
	      Numnodes++;
	      nfree++;
	    }
	}
    }
  last = -1;
  force_update_node_recursive(N, -1, -1);
  if(last >= N) Nodes[last - N].nextnode = -1;
  else Nextnode[last] = -1;
  TreeReconstructFlag = 0;
  return Numnodes;
}

static void force_update_hmax(void)	/* forcetree.c:1661-1705 */
{
  int i, no;
  for(i = 0; i < N; i++)
    if(Type[i] == 0)
      {
	no = Father[i];
	while(no >= 0)
	  {
	    if(Hsml[i] > Nodes[no - N].hmax)
	      {
		Nodes[no - N].hmax = Hsml[i];
		no = Nodes[no - N].father;
	      }
	    else
	      break;
	  }
      }
}

/* ---------------------------------------------------------------- gravity ---- */
static int force_treeevaluate(int target)	/* forcetree.c:1797-2317, mode 0 */
{
  int no, ninteractions = 0, ptype = Type[target];
  double r2, dx, dy, dz, mass, r, fac, u, h = 0, h_inv, h3_inv, acc_x = 0, acc_y = 0, acc_z = 0;
  double pos_x = Pos[3 * target], pos_y = Pos[3 * target + 1], pos_z = Pos[3 * target + 2];
  double aold = ErrTolForceAcc * OldAcc[target], soft = Hsml[target];
  struct node *nop = 0;
  no = N;
  while(no >= 0)
    {
      if(no < N)
	{
	  dx = Pos[3 * no] - pos_x; dy = Pos[3 * no + 1] - pos_y; dz = Pos[3 * no + 2] - pos_z;
	  mass = Mass[no];
	}
      else
	{
	  nop = &Nodes[no - N];
	  mass = nop->mass;
	  if(!nop->multiple && mass)
	    {
	      no = nop->nextnode;
	      continue;
	    }
	  dx = nop->s[0] - pos_x; dy = nop->s[1] - pos_y; dz = nop->s[2] - pos_z;
	}
      r2 = dx * dx + dy * dy + dz * dz;
      if(no < N)
	{
	  h = (ptype == 0) ? soft : ForceSoftening[ptype];	/* :2039-2062 */
	  if(Type[no] == 0)
	    {
	      if(h < Hsml[no]) h = Hsml[no];
	    }
	  else if(h < ForceSoftening[Type[no]])
	    h = ForceSoftening[Type[no]];
	  no = Nextnode[no];
	}
      else
	{
	  if(ErrTolTheta)
	    {
	      if(nop->len * nop->len > r2 * ErrTolTheta * ErrTolTheta)	/* :2076 */
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  else
	    {
	      if(mass * nop->len * nop->len > r2 * r2 * aold)	/* :2085 */
		{
		  no = nop->nextnode;
		  continue;
		}
	      if(fabs(nop->center[0] - pos_x) < 0.60 * nop->len)	/* :2094-2105 */
		if(fabs(nop->center[1] - pos_y) < 0.60 * nop->len)
		  if(fabs(nop->center[2] - pos_z) < 0.60 * nop->len)
		    {
		      no = nop->nextnode;
		      continue;
		    }
	    }
	  h = (ptype == 0) ? soft : ForceSoftening[ptype];	/* :2117-2137 */
	  if(h < nop->maxsoft)
	    {
	      h = nop->maxsoft;
	      if(r2 < h * h)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  no = nop->sibling;
	}
      r = sqrt(r2);
      if(r >= h)
	fac = mass / (r2 * r);
      else
	{
	  h_inv = 1.0 / h;
	  h3_inv = h_inv * h_inv * h_inv;
	  u = r * h_inv;
	  if(u < 0.5)
	    fac = mass * h3_inv * (10.666666666667 + u * u * (32.0 * u - 38.4));
	  else
	    fac = mass * h3_inv * (21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u * u * u - 0.066666666667 / (u * u * u));
	}
      acc_x += dx * fac; acc_y += dy * fac; acc_z += dz * fac;
      if(mass > 0) ninteractions++;
    }
  GravAccel[3 * target] = acc_x; GravAccel[3 * target + 1] = acc_y; GravAccel[3 * target + 2] = acc_z;
  GravCost[target] = ninteractions;
  return ninteractions;
}

static void gravity_tree(void)	/* gravtree.c:27-828 */
{
  int i, j;
  if(TreeReconstructFlag) force_treebuild();
  for(i = 0; i < N; i++) force_treeevaluate(i);
  for(i = 0; i < N; i++)
    {
      double ax = GravAccel[3 * i], ay = GravAccel[3 * i + 1], az = GravAccel[3 * i + 2];
      OldAcc[i] = sqrt(ax * ax + ay * ay + az * az);	/* :392 */
    }
  if(TypeOfOpeningCriterion == 1) ErrTolTheta = 0;	/* :396-397 */
  for(i = 0; i < N; i++)
    for(j = 0; j < 3; j++) GravAccel[3 * i + j] *= G;	/* :400-403 */
}

/* ---------------------------------------------------------------- potential ---- */
static void force_treeevaluate_potential(int target)	/* forcetree.c:3217-3543, mode 0 */
{
  int no, ptype = Type[target];
  double r2, dx, dy, dz, mass, r, u, h = 0, h_inv, wp, pot = 0;
  double pos_x = Pos[3 * target], pos_y = Pos[3 * target + 1], pos_z = Pos[3 * target + 2];
  double aold = ErrTolForceAcc * OldAcc[target], soft = Hsml[target];
  struct node *nop = 0;
  no = N;
  while(no >= 0)
    {
      if(no < N)
	{
	  dx = Pos[3 * no] - pos_x; dy = Pos[3 * no + 1] - pos_y; dz = Pos[3 * no + 2] - pos_z;
	  mass = Mass[no];
	}
      else
	{
	  nop = &Nodes[no - N];
	  mass = nop->mass;
	  if(!nop->multiple && mass)
	    {
	      no = nop->nextnode;
	      continue;
	    }
	  dx = nop->s[0] - pos_x; dy = nop->s[1] - pos_y; dz = nop->s[2] - pos_z;
	}
      r2 = dx * dx + dy * dy + dz * dz;
      if(no < N)
	{
	  h = (ptype == 0) ? soft : ForceSoftening[ptype];
	  if(Type[no] == 0)
	    {
	      if(h < Hsml[no]) h = Hsml[no];
	    }
	  else if(h < ForceSoftening[Type[no]])
	    h = ForceSoftening[Type[no]];
	  no = Nextnode[no];
	}
      else
	{
	  if(ErrTolTheta)
	    {
	      if(nop->len * nop->len > r2 * ErrTolTheta * ErrTolTheta)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  else
	    {
	      if(mass * nop->len * nop->len > r2 * r2 * aold)
		{
		  no = nop->nextnode;
		  continue;
		}
	      if(fabs(nop->center[0] - pos_x) < 0.60 * nop->len)
		if(fabs(nop->center[1] - pos_y) < 0.60 * nop->len)
		  if(fabs(nop->center[2] - pos_z) < 0.60 * nop->len)
		    {
		      no = nop->nextnode;
		      continue;
		    }
	    }
	  h = (ptype == 0) ? soft : ForceSoftening[ptype];
	  if(h < nop->maxsoft)
	    {
	      h = nop->maxsoft;
	      if(r2 < h * h)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  no = nop->sibling;
	}
      r = sqrt(r2);
      if(r >= h)
	pot += (-mass / r);
      else
	{
	  h_inv = 1.0 / h;
	  u = r * h_inv;
	  if(u < 0.5)
	    wp = -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
	  else
	    wp = -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
	  pot += (mass * h_inv * wp);
	}
    }
  Potential[target] = pot;
}

static void compute_potential(void)	/* potential.c:22-250 */
{
  int i;
  if(TreeReconstructFlag) force_treebuild();
  for(i = 0; i < N; i++) force_treeevaluate_potential(i);
  for(i = 0; i < N; i++)
    {
      Potential[i] += Mass[i] / SofteningTable[Type[i]];	/* remove self-potential, potential.c:176 */
      Potential[i] *= G;	/* :184 */
    }
}

/* ---------------------------------------------------------------- neighbour search ---- */
static int *Ngblist;
static int ngb_treefind_variable(const double *c, double hsml)	/* ngb.c:169-297 */
{
  int numngb = 0, no = N, p;
  double dx, dy, dz, dist;
  while(no >= 0)
    {
      if(no < N)
	{
	  p = no;
	  no = Nextnode[no];
	  if(Type[p] > 0) continue;
	  dist = hsml;
	  dx = fabs(Pos[3 * p] - c[0]); if(dx > dist) continue;
	  dy = fabs(Pos[3 * p + 1] - c[1]); if(dy > dist) continue;
	  dz = fabs(Pos[3 * p + 2] - c[2]); if(dz > dist) continue;
	  if(dx * dx + dy * dy + dz * dz > dist * dist) continue;
	  Ngblist[numngb++] = p;
	}
      else
	{
	  struct node *cur = &Nodes[no - N];
	  no = cur->sibling;
	  dist = hsml + 0.5 * cur->len;
	  dx = fabs(cur->center[0] - c[0]); if(dx > dist) continue;
	  dy = fabs(cur->center[1] - c[1]); if(dy > dist) continue;
	  dz = fabs(cur->center[2] - c[2]); if(dz > dist) continue;
	  dist += 0.366025403785 * cur->len;
	  if(dx * dx + dy * dy + dz * dz > dist * dist) continue;
	  no = cur->nextnode;
	}
    }
  return numngb;
}

static int ngb_treefind_pairs(const double *c, double hsml)	/* ngb.c:32-160 */
{
  int numngb = 0, no = N, p;
  double dx, dy, dz, dist;
  while(no >= 0)
    {
      if(no < N)
	{
	  p = no;
	  no = Nextnode[no];
	  if(Type[p] > 0) continue;
	  dist = (Hsml[p] > hsml) ? Hsml[p] : hsml;
	  dx = fabs(Pos[3 * p] - c[0]); if(dx > dist) continue;
	  dy = fabs(Pos[3 * p + 1] - c[1]); if(dy > dist) continue;
	  dz = fabs(Pos[3 * p + 2] - c[2]); if(dz > dist) continue;
	  if(dx * dx + dy * dy + dz * dz > dist * dist) continue;
	  Ngblist[numngb++] = p;
	}
      else
	{
	  struct node *cur = &Nodes[no - N];
	  dist = ((cur->hmax > hsml) ? cur->hmax : hsml) + 0.5 * cur->len;
	  no = cur->sibling;
	  dx = fabs(cur->center[0] - c[0]); if(dx > dist) continue;
	  dy = fabs(cur->center[1] - c[1]); if(dy > dist) continue;
	  dz = fabs(cur->center[2] - c[2]); if(dz > dist) continue;
	  dist += 0.366025403785 * cur->len;
	  if(dx * dx + dy * dy + dz * dz > dist * dist) continue;
	  no = cur->nextnode;
	}
    }
  return numngb;
}

/* ---------------------------------------------------------------- density ---- */
static int *DensPairs, *HydroPairs;
static void density_evaluate(int target)	/* density.c:711-1029, mode 0 */
{
  int j, n, numngb_inbox, numngb = 0;
  double h = Hsml[target], h2, hinv, hinv3, hinv4, rho = 0, wk, dwk, dx, dy, dz, r, r2, u, mass_j, dvx, dvy, dvz, fac,
    weighted_numngb = 0, dhsmlrho = 0, gasvel[3] = { 0, 0, 0 }, divv = 0, rotv[3] = { 0, 0, 0 }, smoothentr = 0;
  const double *pos = &Pos[3 * target];
  static const double veldummy[3] = { 0, 0, 0 };
  const double *vel = (Type[target] == 0) ? &Vel[3 * target] : veldummy;
  h2 = h * h; hinv = 1.0 / h; hinv3 = hinv * hinv * hinv; hinv4 = hinv3 * hinv;
  numngb_inbox = ngb_treefind_variable(pos, h);
  for(n = 0; n < numngb_inbox; n++)
    {
      j = Ngblist[n];
      if(Mass[j] == 0) continue;
      dx = pos[0] - Pos[3 * j]; dy = pos[1] - Pos[3 * j + 1]; dz = pos[2] - Pos[3 * j + 2];
      r2 = dx * dx + dy * dy + dz * dz;
      if(r2 < h2)
	{
	  numngb++;
	  r = sqrt(r2);
	  u = r * hinv;
	  if(u < 0.5)
	    {
	      wk = hinv3 * (KC1 + KC2 * (u - 1) * u * u);
	      dwk = hinv4 * u * (KC3 * u - KC4);
	    }
	  else
	    {
	      wk = hinv3 * KC5 * (1.0 - u) * (1.0 - u) * (1.0 - u);
	      dwk = hinv4 * KC6 * (1.0 - u) * (1.0 - u);
	    }
	  mass_j = Mass[j];
	  rho += mass_j * wk;
	  weighted_numngb += NORM * wk / hinv3;
	  dhsmlrho += -mass_j * (3 * hinv * wk + u * dwk);
	  gasvel[0] += mass_j * wk * Vel[3 * j]; gasvel[1] += mass_j * wk * Vel[3 * j + 1]; gasvel[2] += mass_j * wk * Vel[3 * j + 2];
	  smoothentr += mass_j * wk * Entropy[j];
	  if(r > 0)
	    {
	      fac = mass_j * dwk / r;
	      dvx = vel[0] - Vel[3 * j]; dvy = vel[1] - Vel[3 * j + 1]; dvz = vel[2] - Vel[3 * j + 2];
	      divv += -fac * (dx * dvx + dy * dvy + dz * dvz);
	      rotv[0] += fac * (dz * dvy - dy * dvz);
	      rotv[1] += fac * (dx * dvz - dz * dvx);
	      rotv[2] += fac * (dy * dvx - dx * dvy);
	    }
	}
    }
  NumNgb[target] = weighted_numngb;
  DensPairs[target] = numngb;
  if(Type[target] == 0)
    {
      Density[target] = rho; DhsmlDensityFactor[target] = dhsmlrho; DivVel[target] = divv;
      CurlVel[target] = sqrt(rotv[0] * rotv[0] + rotv[1] * rotv[1] + rotv[2] * rotv[2]);	/* |rot|, divided by rho in density() */
    }
  SinkDensity[target] = rho; SinkEntropy[target] = smoothentr;
  SinkGasVel[3 * target] = gasvel[0]; SinkGasVel[3 * target + 1] = gasvel[1]; SinkGasVel[3 * target + 2] = gasvel[2];
}

static int density_isactive(int n)	/* density.c:1035-1057 */
{
  if(TimeBin[n] < 0) return 0;
  return Type[n] == 5 || Type[n] == 2 || Type[n] == 0;
}

static int density_passes;
static void density(void)	/* density.c:89-704 */
{
  double *Left = calloc(N, sizeof(double)), *Right = calloc(N, sizeof(double)), desnumngb, fac, dt_entr;
  int i, npleft, iter = 0, dt_step;
  density_passes = 0;
  do
    {
      for(i = 0; i < N; i++)
	if(density_isactive(i)) density_evaluate(i);
      density_passes++;
      for(i = 0, npleft = 0; i < N; i++)
	{
	  if(!density_isactive(i)) continue;
	  if(Type[i] == 0)
	    {
	      if(Density[i] > 0)
		{
		  DhsmlDensityFactor[i] *= Hsml[i] / (3 * Density[i]);	/* :442-446 */
		  if(DhsmlDensityFactor[i] > -0.9) DhsmlDensityFactor[i] = 1 / (1 + DhsmlDensityFactor[i]);
		  else DhsmlDensityFactor[i] = 1;
		  CurlVel[i] = CurlVel[i] / Density[i];	/* :448-450 */
		  DivVel[i] /= Density[i];
		}
	      dt_step = TimeBin[i] ? (1 << TimeBin[i]) : 0;
	      dt_entr = (Ti_Current - (0 + dt_step / 2)) * Timebase_interval;	/* Ti_begstep = 0 in harness inputs */
	      Pressure[i] = (Entropy[i] + DtEntropyIn[i] * dt_entr) * pow(Density[i], GAMMA);	/* :490-496 */
	    }
	  else if(SinkDensity[i] > 0)	/* :520-546 */
	    {
	      SinkEntropy[i] /= SinkDensity[i];
	      SinkGasVel[3 * i] /= SinkDensity[i]; SinkGasVel[3 * i + 1] /= SinkDensity[i]; SinkGasVel[3 * i + 2] /= SinkDensity[i];
	    }
	  desnumngb = (Type[i] == 5) ? DesNumNgb * BlackHoleNgbFactor : DesNumNgb;
	  if(NumNgb[i] < (desnumngb - MaxNumNgbDeviation) || (NumNgb[i] > (desnumngb + MaxNumNgbDeviation) && Hsml[i] > (1.01 * MinGasHsml)))
	    {
	      npleft++;
	      if(Left[i] > 0 && Right[i] > 0)
		if((Right[i] - Left[i]) < 1.0e-3 * Left[i])
		  {
		    npleft--;
		    TimeBin[i] = -TimeBin[i] - 1;
		    continue;
		  }
	      if(NumNgb[i] < (desnumngb - MaxNumNgbDeviation))
		Left[i] = (Hsml[i] > Left[i]) ? Hsml[i] : Left[i];
	      else
		{
		  if(Right[i] != 0)
		    {
		      if(Hsml[i] < Right[i]) Right[i] = Hsml[i];
		    }
		  else
		    Right[i] = Hsml[i];
		}
	      if(Right[i] > 0 && Left[i] > 0)
		Hsml[i] = pow(0.5 * (pow(Left[i], 3) + pow(Right[i], 3)), 1.0 / 3);
	      else
		{
		  if(Right[i] == 0 && Left[i] == 0) { fprintf(stderr, "endrun(8188)\n"); exit(8); }
		  if(Right[i] == 0 && Left[i] > 0)
		    {
		      if(Type[i] == 0 && fabs(NumNgb[i] - desnumngb) < 0.5 * desnumngb)
			{
			  fac = 1 - (NumNgb[i] - desnumngb) / (3 * NumNgb[i]) * DhsmlDensityFactor[i];
			  if(fac < 1.26) Hsml[i] *= fac;
			  else Hsml[i] *= 1.26;
			}
		      else
			Hsml[i] *= 1.26;
		    }
		  if(Right[i] > 0 && Left[i] == 0)
		    {
		      if(Type[i] == 0 && fabs(NumNgb[i] - desnumngb) < 0.5 * desnumngb)
			{
			  fac = 1 - (NumNgb[i] - desnumngb) / (3 * NumNgb[i]) * DhsmlDensityFactor[i];
			  if(fac > 1 / 1.26) Hsml[i] *= fac;
			  else Hsml[i] /= 1.26;
			}
		      else
			Hsml[i] /= 1.26;
		    }
		}
	      if(Hsml[i] < MinGasHsml) Hsml[i] = MinGasHsml;
	    }
	  else
	    TimeBin[i] = -TimeBin[i] - 1;
	}
      if(npleft > 0)
	{
	  iter++;
	  if(iter > MAXITER) { fprintf(stderr, "endrun(1155)\n"); exit(9); }
	}
    }
  while(npleft > 0);
  for(i = 0; i < N; i++)
    if(TimeBin[i] < 0) TimeBin[i] = -TimeBin[i] - 1;
  free(Left); free(Right);
}

/* ---------------------------------------------------------------- hydro ---- */
static void hydro_evaluate(int target)	/* hydra.c:822-1995, live lines (non-comoving) */
{
  int j, n, numngb, timestep, npairs = 0, ts_j;
  const double *pos = &Pos[3 * target], *vel = &Vel[3 * target];
  double h_i = Hsml[target], mass = Mass[target], rho = Density[target], pressure = Pressure[target];
  double dx, dy, dz, dvx, dvy, dvz, h_i2, hinv, hinv4, p_over_rho2_i, p_over_rho2_j, soundspeed_i, soundspeed_j, hfc, dwk_i,
    vdotr, vdotr2, visc, mu_ij, rho_ij, vsig, h_j, dwk_j, r, r2, u, hfc_visc, f1, f2, dt, acc[3] = { 0, 0, 0 }, dtEntropy = 0,
    maxSignalVel = 0, lim;
  timestep = TimeBin[target] ? (1 << TimeBin[target]) : 0;
  soundspeed_i = sqrt(GAMMA * pressure / rho);
  f1 = fabs(DivVel[target]) / (fabs(DivVel[target]) + CurlVel[target] + 0.0001 * soundspeed_i / Hsml[target] / 1.0);
  p_over_rho2_i = pressure / (rho * rho);
  p_over_rho2_i *= DhsmlDensityFactor[target];
  h_i2 = h_i * h_i;
  numngb = ngb_treefind_pairs(pos, h_i);
  for(n = 0; n < numngb; n++)
    {
      j = Ngblist[n];
      if(Mass[j] == 0) continue;
      dx = pos[0] - Pos[3 * j]; dy = pos[1] - Pos[3 * j + 1]; dz = pos[2] - Pos[3 * j + 2];
      r2 = dx * dx + dy * dy + dz * dz;
      h_j = Hsml[j];
      if(r2 < h_i2 || r2 < h_j * h_j)
	{
	  r = sqrt(r2);
	  if(r > 0)
	    {
	      npairs++;
	      p_over_rho2_j = Pressure[j] / (Density[j] * Density[j]);
	      soundspeed_j = sqrt(GAMMA * p_over_rho2_j * Density[j]);
	      dvx = vel[0] - Vel[3 * j]; dvy = vel[1] - Vel[3 * j + 1]; dvz = vel[2] - Vel[3 * j + 2];
	      vdotr = dx * dvx + dy * dvy + dz * dvz;
	      vdotr2 = vdotr;
	      if(r2 < h_i2)
		{
		  hinv = 1.0 / h_i; hinv4 = hinv * hinv * hinv * hinv; u = r * hinv;
		  dwk_i = (u < 0.5) ? hinv4 * u * (KC3 * u - KC4) : hinv4 * KC6 * (1.0 - u) * (1.0 - u);
		}
	      else
		dwk_i = 0;
	      if(r2 < h_j * h_j)
		{
		  hinv = 1.0 / h_j; hinv4 = hinv * hinv * hinv * hinv; u = r * hinv;
		  dwk_j = (u < 0.5) ? hinv4 * u * (KC3 * u - KC4) : hinv4 * KC6 * (1.0 - u) * (1.0 - u);
		}
	      else
		dwk_j = 0;
	      vsig = soundspeed_i + soundspeed_j;
	      if(vsig > maxSignalVel) maxSignalVel = vsig;
	      if(vdotr2 < 0)
		{
		  mu_ij = 1.0 * vdotr2 / r;
		  vsig -= 3 * mu_ij;
		  if(vsig > maxSignalVel) maxSignalVel = vsig;
		  rho_ij = 0.5 * (rho + Density[j]);
		  f2 = fabs(DivVel[j]) / (fabs(DivVel[j]) + CurlVel[j] + 0.0001 * soundspeed_j / 1.0 / Hsml[j]);
		  visc = 0.25 * ArtBulkViscConst * vsig * (-mu_ij) / rho_ij * (f1 + f2);
		  ts_j = TimeBin[j] ? (1 << TimeBin[j]) : 0;
		  dt = 2 * ((timestep > ts_j) ? timestep : ts_j) * Timebase_interval;
		  if(dt > 0 && (dwk_i + dwk_j) < 0)
		    {
		      lim = 0.5 * 1.0 * vdotr2 / (0.5 * (mass + Mass[j]) * (dwk_i + dwk_j) * r * dt);
		      visc = (visc < lim) ? visc : lim;
		    }
		}
	      else
		visc = 0;
	      p_over_rho2_j *= DhsmlDensityFactor[j];
	      hfc_visc = 0.5 * Mass[j] * visc * (dwk_i + dwk_j) / r;
	      hfc = hfc_visc + Mass[j] * (p_over_rho2_i * dwk_i + p_over_rho2_j * dwk_j) / r;
	      acc[0] += (-hfc * dx); acc[1] += (-hfc * dy); acc[2] += (-hfc * dz);
	      dtEntropy += (0.5 * hfc_visc * vdotr2);
	    }
	}
    }
  HydroAccel[3 * target] = acc[0]; HydroAccel[3 * target + 1] = acc[1]; HydroAccel[3 * target + 2] = acc[2];
  DtEntropy[target] = dtEntropy;
  MaxSignalVel[target] = maxSignalVel;
  HydroPairs[target] = npairs;
}

static void hydro_force(void)	/* hydra.c:145-813 */
{
  int i;
  for(i = 0; i < N; i++)
    if(Type[i] == 0) hydro_evaluate(i);
  for(i = 0; i < N; i++)
    if(Type[i] == 0) DtEntropy[i] *= GAMMA_MINUS1 / (1.0 * pow(Density[i], GAMMA_MINUS1));	/* :583 */
}

static void initial_hsml(void)	/* init.c:719-745 */
{
  int i, no, p;
  for(i = 0; i < N; i++)
    {
      if(Type[i] == 0)
	{
	  no = Father[i];
	  while(10 * DesNumNgb * Mass[i] > Nodes[no - N].mass)
	    {
	      p = Nodes[no - N].father;
	      if(p < 0) break;
	      no = p;
	    }
	  Hsml[i] = pow(3.0 / (4 * M_PI) * DesNumNgb * Mass[i] / Nodes[no - N].mass, 1.0 / 3) * Nodes[no - N].len;
	  if(Hsml[i] > 10.0 * SofteningTable[0]) Hsml[i] = SofteningTable[0];
	}
      if(Type[i] == 5) Hsml[i] = SofteningTable[5];
      if(Type[i] == 2) Hsml[i] = SofteningTable[2];
    }
}

/* ---------------------------------------------------------------- output ---- */
static FILE *fout;
static void put(const char *prefix, const char *name, char kind, long long n, int ncomp, const void *data)
{
  char tag[48];
  int esz = (kind == 'd' || kind == 'q') ? 8 : 4;
  memset(tag, 0, sizeof(tag));
  snprintf(tag, sizeof(tag), "%s%s", prefix, name);
  fwrite(tag, 1, 48, fout); fwrite(&kind, 1, 1, fout); fwrite(&ncomp, 4, 1, fout); fwrite(&n, 8, 1, fout);
  fwrite(data, esz, (size_t) n * ncomp, fout);
}

static void dump(const char *prefix, char op)
{
  put(prefix, "id", 'i', N, 1, ID);
  if(op == 'G')
    {
      put(prefix, "GravAccel", 'd', N, 3, GravAccel); put(prefix, "OldAcc", 'd', N, 1, OldAcc); put(prefix, "GravCost", 'd', N, 1, GravCost);
    }
  if(op == 'D' || op == 'S') put(prefix, "Hsml", 'd', N, 1, Hsml);
  if(op == 'D')
    {
      int i;
      double *sd = calloc(N, 8), *se = calloc(N, 8), *sv = calloc(3 * (size_t) N, 8);
      for(i = 0; i < N; i++)
	if(Type[i] == 5 || Type[i] == 2)
	  {
	    sd[i] = SinkDensity[i]; se[i] = SinkEntropy[i];
	    sv[3 * i] = SinkGasVel[3 * i]; sv[3 * i + 1] = SinkGasVel[3 * i + 1]; sv[3 * i + 2] = SinkGasVel[3 * i + 2];
	  }
      put(prefix, "NumNgb", 'd', N, 1, NumNgb); put(prefix, "Density", 'd', Ngas, 1, Density);
      put(prefix, "DhsmlDensityFactor", 'd', Ngas, 1, DhsmlDensityFactor); put(prefix, "DivVel", 'd', Ngas, 1, DivVel);
      put(prefix, "CurlVel", 'd', Ngas, 1, CurlVel); put(prefix, "Pressure", 'd', Ngas, 1, Pressure);
      put(prefix, "SinkDensity", 'd', N, 1, sd); put(prefix, "SinkEntropy", 'd', N, 1, se); put(prefix, "SinkGasVel", 'd', N, 3, sv);
      free(sd); free(se); free(sv);
    }
  if(op == 'H')
    {
      put(prefix, "HydroAccel", 'd', Ngas, 3, HydroAccel); put(prefix, "DtEntropy", 'd', Ngas, 1, DtEntropy);
      put(prefix, "MaxSignalVel", 'd', Ngas, 1, MaxSignalVel);
    }
}

static void dump_nodes(const char *prefix)
{
  long long k;
  double *buf = malloc(sizeof(double) * 11 * (size_t) Numnodes), dom[8];
  int *fl = malloc(sizeof(int) * (size_t) Numnodes);
  for(k = 0; k < Numnodes; k++)
    {
      struct node *nd = &Nodes[k];
      buf[11 * k] = nd->len; buf[11 * k + 1] = nd->center[0]; buf[11 * k + 2] = nd->center[1]; buf[11 * k + 3] = nd->center[2];
      buf[11 * k + 4] = nd->mass; buf[11 * k + 5] = nd->s[0]; buf[11 * k + 6] = nd->s[1]; buf[11 * k + 7] = nd->s[2];
      buf[11 * k + 8] = nd->maxsoft; buf[11 * k + 9] = nd->hmax; buf[11 * k + 10] = nd->vmax;
      fl[k] = nd->multiple ? 128 : 0;	/* BITFLAG_MULTIPLEPARTICLES = 7 */
    }
  put(prefix, "nodes", 'd', Numnodes, 11, buf); put(prefix, "nodeflags", 'i', Numnodes, 1, fl);
  dom[0] = DomainLen; dom[1] = DomainFac;
  memcpy(dom + 2, DomainCorner, 24); memcpy(dom + 5, DomainCenter, 24);
  put(prefix, "domain", 'd', 8, 1, dom);
  free(buf); free(fl);
}

static void pair_counts(const char *prefix)
{
  /* the counts of the LAST density()/hydro_force() evaluation at the current Hsml are what the reference
   * harness recomputes; re-evaluate both here so the op is self-contained */
  int i;
  int *cd = calloc(N, sizeof(int)), *ch = calloc(N, sizeof(int));
  for(i = 0; i < N; i++)
    {
      int k, nn;
      double h = Hsml[i], h2 = h * h;
      if(!(Type[i] == 0 || Type[i] == 5 || Type[i] == 2)) continue;
      nn = ngb_treefind_variable(&Pos[3 * i], h);
      for(k = 0; k < nn; k++)
	{
	  int j = Ngblist[k];
	  double dx = Pos[3 * i] - Pos[3 * j], dy = Pos[3 * i + 1] - Pos[3 * j + 1], dz = Pos[3 * i + 2] - Pos[3 * j + 2];
	  if(Mass[j] == 0) continue;
	  if(dx * dx + dy * dy + dz * dz < h2) cd[i]++;
	}
      if(Type[i] != 0) continue;
      nn = ngb_treefind_pairs(&Pos[3 * i], h);
      for(k = 0; k < nn; k++)
	{
	  int j = Ngblist[k];
	  double dx = Pos[3 * i] - Pos[3 * j], dy = Pos[3 * i + 1] - Pos[3 * j + 1], dz = Pos[3 * i + 2] - Pos[3 * j + 2], r2;
	  if(Mass[j] == 0) continue;
	  r2 = dx * dx + dy * dy + dz * dz;
	  if((r2 < h2 || r2 < Hsml[j] * Hsml[j]) && sqrt(r2) > 0) ch[i]++;
	}
    }
  put(prefix, "id", 'i', N, 1, ID);
  put(prefix, "DensPairs", 'i', N, 1, cd); put(prefix, "HydroPairs", 'i', N, 1, ch);
  free(cd); free(ch);
}

/* also used by tests: dump keys of every particle */
static void dump_keys(const char *prefix)
{
  long long *ph = malloc(8 * (size_t) N), *mo = malloc(8 * (size_t) N);
  int i;
  for(i = 0; i < N; i++)
    {
      int x = (int) ((Pos[3 * i] - DomainCorner[0]) * DomainFac), y = (int) ((Pos[3 * i + 1] - DomainCorner[1]) * DomainFac),
	z = (int) ((Pos[3 * i + 2] - DomainCorner[2]) * DomainFac);
      ph[i] = (long long) peano_hilbert_key(x, y, z, BITS);
      mo[i] = (long long) morton_key(x, y, z, BITS);
    }
  put(prefix, "id", 'i', N, 1, ID);
  put(prefix, "phkey", 'q', N, 1, ph); put(prefix, "mortonkey", 'q', N, 1, mo);
  free(ph); free(mo);
}

int main(int argc, char **argv)
{
  FILE *fin;
  char magic[4], *ops, *tok;
  int version, i, opidx = 0;
  long long n8, ng8;
  double t0, dt, interactions;
  if(argc < 4) { fprintf(stderr, "usage: %s in out \"ops\"\n", argv[0]); return 2; }
  fin = fopen(argv[1], "rb");
  if(!fin) { perror(argv[1]); return 2; }
  if(fread(magic, 1, 4, fin) != 4 || memcmp(magic, "GLIN", 4)) return 2;
  if(fread(&version, 4, 1, fin) != 1 || fread(&n8, 8, 1, fin) != 1 || fread(&ng8, 8, 1, fin) != 1) return 2;
  if(fread(par, 8, NPAR, fin) != NPAR) return 2;
  N = (int) n8; Ngas = (int) ng8;
#define RD(ptr, T, cnt) do { ptr = malloc(sizeof(T) * (size_t) (cnt) + 8); if(fread(ptr, sizeof(T), (size_t) (cnt), fin) != (size_t) (cnt)) return 2; } while(0)
  RD(Pos, double, 3 * (size_t) N); RD(Vel, double, 3 * (size_t) N); RD(Mass, double, N); RD(Type, int, N); RD(Hsml, double, N);
  RD(ID, int, N); RD(TimeBin, int, N); RD(OldAcc, double, N); RD(Entropy, double, Ngas); RD(DtEntropyIn, double, Ngas);
  fclose(fin);
  Entropy = realloc(Entropy, 8 * (size_t) N + 8); DtEntropyIn = realloc(DtEntropyIn, 8 * (size_t) N + 8);
  for(i = Ngas; i < N; i++) Entropy[i] = DtEntropyIn[i] = 0;
  ErrTolTheta = par[P_THETA]; ErrTolForceAcc = par[P_ERRTOLFORCEACC]; TypeOfOpeningCriterion = (int) par[P_OPENCRIT];
  DesNumNgb = (int) par[P_DESNUMNGB]; MaxNumNgbDeviation = par[P_MAXNGBDEV]; ArtBulkViscConst = par[P_ARTVISC]; G = par[P_G];
  for(i = 0; i < 6; i++)
    {
      SofteningTable[i] = par[P_SOFT0 + i];
      ForceSoftening[i] = 2.8 * SofteningTable[i];	/* gravtree.c:881-882 */
    }
  MinGasHsml = par[P_MINGASHSMLFRAC] * ForceSoftening[0];	/* gravtree.c:884 */
  Timebase_interval = par[P_TIMEBASE_INTERVAL]; Ti_Current = (int) par[P_TI_CURRENT]; BlackHoleNgbFactor = par[P_BHNGBFACTOR];
#define AL(ptr, cnt) ptr = calloc((size_t) (cnt) + 1, sizeof(*ptr))
  AL(Potential, N);
  AL(GravAccel, 3 * (size_t) N); AL(GravCost, N); AL(NumNgb, N); AL(Density, N); AL(DhsmlDensityFactor, N); AL(DivVel, N);
  AL(CurlVel, N); AL(Pressure, N); AL(SinkDensity, N); AL(SinkEntropy, N); AL(SinkGasVel, 3 * (size_t) N); AL(HydroAccel, 3 * (size_t) N);
  AL(DtEntropy, N); AL(MaxSignalVel, N); AL(Nextnode, N); AL(Father, N); AL(MortonKey, N); AL(Ngblist, N); AL(DensPairs, N);
  AL(HydroPairs, N);
  MaxNodes = 2 * N + 64;
  Nodes = calloc((size_t) MaxNodes, sizeof(struct node));
  for(i = 0; i < Ngas; i++) Density[i] = -1;
  fout = fopen(argv[2], "wb");
  if(!fout) { perror(argv[2]); return 2; }
  fwrite("GLOU", 1, 4, fout);
  t0 = now();
  domain_findExtent();
  dt = now() - t0;
  put("", "time_domain", 'd', 1, 1, &dt);
  ops = strdup(argv[3]);
  for(tok = strtok(ops, " "); tok; tok = strtok(NULL, " "), opidx++)
    {
      char prefix[32];
      const char *seq = (tok[0] == 'A') ? "RGDUH" : tok;
      for(; *seq; seq++)
	{
	  char op = *seq;
	  snprintf(prefix, sizeof(prefix), "%d_%c_", opidx, op);
	  t0 = now();
	  switch (op)
	    {
	    case 'T': force_treebuild(); break;
	    case 'S':
	      initial_hsml();
	      for(i = 0; i < N; i++)	/* same input convention as ref_harness.c: OldAcc < 0 = swallowed after this step */
		if(OldAcc[i] < 0)
		  {
		    Mass[i] = 0;
		    OldAcc[i] = 0;
		  }
	      break;
	    case 'D': density(); break;
	    case 'R': TreeReconstructFlag = 1; break;
	    case 'G': gravity_tree(); break;
	    case 'U': force_update_hmax(); break;
	    case 'H': hydro_force(); break;
	    case 'N': dump_nodes(prefix); break;
	    case 'C': pair_counts(prefix); break;
	    case 'K': dump_keys(prefix); break;
	    case 'P': compute_potential(); break;
	    default: fprintf(stderr, "unknown op %c\n", op); return 2;
	    }
	  dt = now() - t0;
	  if(strchr("TDGUHP", op)) put(prefix, "time", 'd', 1, 1, &dt);
	  if(op == 'P')
	    {
	      put(prefix, "id", 'i', N, 1, ID);
	      put(prefix, "Potential", 'd', N, 1, Potential);
	    }
	  if(strchr("GDHS", op)) dump(prefix, op);
	  if(op == 'D') { double pp = density_passes; put(prefix, "passes", 'd', 1, 1, &pp); }
	  if(op == 'G')
	    {
	      interactions = 0;
	      for(i = 0; i < N; i++) interactions += GravCost[i];
	      put(prefix, "interactions", 'd', 1, 1, &interactions);
	      put(prefix, "theta_after", 'd', 1, 1, &ErrTolTheta);
	    }
	  if(op == 'T' || op == 'G') put(prefix, "numnodes", 'i', 1, 1, &Numnodes);
	}
    }
  fclose(fout);
  return 0;
}
