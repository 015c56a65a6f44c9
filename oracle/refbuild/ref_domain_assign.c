/* ref_domain_assign.c - the multi-rank half of the reference's domain_decompose() (domain.c:470-487) run on given
 * leaf tables: domain_findSplit_work_balanced (:1075), domain_assign_load_or_work_balanced (:1153),
 * domain_check_memory_bound (:595), domain_findSplit_load_balanced (:1225) - the reference's own functions with NTask
 * set to what the caller asks for.  No second rank is needed: these functions run identically on every rank, on tables
 * that domain_sumCost has already summed over the ranks.  Their tables are file-static in domain.c, so this driver
 * includes the reference source where it lies (it is never copied) and is linked with the other reference objects.
 *
 * TEST INFRASTRUCTURE ONLY (target `domain_assign` -> oracle/_ref/ref_domain_assign), the checker of gl_domain_assign.
 *
 * usage: ref_domain_assign <in> <out>
 *   in : int NTask, ntopleaves, maxLoad, maxLoadsph; float work[ntopleaves]; int count[ntopleaves], countsph[ntopleaves]
 *   out: int status (0 first variant, 1 second, 2 none fits); int DomainTask[ntopleaves];
 *        int DomainStartList[16*NTask], DomainEndList[16*NTask]; int list_load[NTask], list_loadsph[NTask]; double list_work[NTask]
 */
#include "domain.c"

double get_drift_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double get_gravkick_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double get_hydrokick_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double hubble_function(double a) { (void) a; return 0; }
void rearrange_particle_sequence(void) { }
double gsl_rng_uniform(gsl_rng * r) { (void) r; return 0.5; }
void compute_accelerations(int mode) { (void) mode; }

int main(int argc, char **argv)
{
  FILE *f;
  int hdr[4], i, status;
  if(argc < 3) return 2;
  f = fopen(argv[1], "rb");
  if(!f || fread(hdr, 4, 4, f) != 4) return 2;
  MPI_Init(&argc, &argv);
  ThisTask = 0;
  NTask = hdr[0];
  NTopleaves = hdr[1];
  maxLoad = hdr[2];
  maxLoadsph = hdr[3];
  All.PartAllocFactor = 1.0;
  domainWork = malloc(sizeof(float) * NTopleaves);
  domainCount = malloc(sizeof(int) * NTopleaves);
  domainCountSph = malloc(sizeof(int) * NTopleaves);
  if(fread(domainWork, 4, NTopleaves, f) != (size_t) NTopleaves || fread(domainCount, 4, NTopleaves, f) != (size_t) NTopleaves ||
     fread(domainCountSph, 4, NTopleaves, f) != (size_t) NTopleaves)
    return 2;
  fclose(f);
  mymalloc_init();
  DomainStartList = malloc(sizeof(int) * MULTIPLEDOMAINS * NTask);
  DomainEndList = malloc(sizeof(int) * MULTIPLEDOMAINS * NTask);
  DomainTask = malloc(sizeof(int) * NTopleaves);
  list_load = malloc(sizeof(int) * NTask);
  list_loadsph = malloc(sizeof(int) * NTask);
  list_work = malloc(sizeof(double) * NTask);
  list_speedfac = malloc(sizeof(double) * NTask);
  for(i = 0; i < NTask; i++) list_speedfac[i] = 1;
  /* domain.c:470-487 */
  domain_findSplit_work_balanced(MULTIPLEDOMAINS * NTask, NTopleaves);
  domain_assign_load_or_work_balanced(1);
  status = domain_check_memory_bound();
  if(status != 0)
    {
      domain_findSplit_load_balanced(MULTIPLEDOMAINS * NTask, NTopleaves);
      domain_assign_load_or_work_balanced(0);
      status = domain_check_memory_bound() ? 2 : 1;
    }
  f = fopen(argv[2], "wb");
  if(!f) return 2;
  fwrite(&status, 4, 1, f);
  fwrite(DomainTask, 4, NTopleaves, f);
  fwrite(DomainStartList, 4, MULTIPLEDOMAINS * NTask, f);
  fwrite(DomainEndList, 4, MULTIPLEDOMAINS * NTask, f);
  fwrite(list_load, 4, NTask, f);
  fwrite(list_loadsph, 4, NTask, f);
  fwrite(list_work, 8, NTask, f);
  fclose(f);
  return 0;
}
