/*
 * oracle/cpipelines.c -- TEST INFRASTRUCTURE / CPU BASELINE ("port").  Not the product path.
 *
 * Plain-C restatement of how the reference executes TPC-H Q1 and Q6 on the CPU, used (a) as a
 * second checker at sizes where oracle/ops.py (Python ints) is too slow and (b) as the
 * `cpu_baseline` / `--impl reference` arm of bench.py.  The reference's operators for this path
 * are DataFusion 53.1 (third party, not vendored); what is restated is their
 * published batch-at-a-time algorithm as Sail drives it:
 *   - one task per partition (`target_partitions` = cores: crates/sail-session/src/
 *     session_factory/server.rs:211-214), 8192-row batches (application.yaml:247-251)
 *   - FilterExec: predicate -> selection, then `filter` (gather) of the projected columns
 *     (plan: python/pysail/tests/spark/__snapshots__/test_tpch.plan.yaml:18)
 *   - ProjectionExec: `price * (1 - disc)` as Decimal128(32,4) i128 multiply (:15)
 *   - AggregateExec(Partial): group ids through a hash table over the key bytes, then per-aggregate
 *     accumulate loops over (values, group ids) (:14); FinalPartitioned merge of the partials (:12)
 * Arrow layouts are consumed in place: Decimal128 = 16-byte LE, Utf8View = 16-byte views,
 * Date32 = int32 days.  Checked against oracle/ops.py and the golden Q1/Q6 snapshot in
 * tests/test_oracle_golden.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;
#define BATCH 8192
#define MAXG 64

typedef struct {
  /* inputs */
  const i128 *qty, *price, *disc, *tax;
  const uint8_t *rflag, *lstatus;     /* 16-byte views */
  const int32_t* shipdate;
  int64_t row0, row1;
  int32_t cutoff;
  /* per-partition partial state */
  int n_groups;
  uint8_t key[MAXG][32];
  i128 sum_qty[MAXG], sum_price[MAXG], sum_disc_price[MAXG], sum_charge[MAXG], sum_disc[MAXG];
  int64_t count[MAXG];
} q1_part;

static int find_group(q1_part* p, const uint8_t* a, const uint8_t* b) {
  for (int g = 0; g < p->n_groups; ++g)
    if (!memcmp(p->key[g], a, 16) && !memcmp(p->key[g] + 16, b, 16)) return g;
  int g = p->n_groups++;
  memcpy(p->key[g], a, 16);
  memcpy(p->key[g] + 16, b, 16);
  p->sum_qty[g] = p->sum_price[g] = p->sum_disc_price[g] = p->sum_charge[g] = p->sum_disc[g] = 0;
  p->count[g] = 0;
  return g;
}

static void* q1_task(void* arg) {
  q1_part* p = (q1_part*)arg;
  static const i128 ONE = 100; /* Decimal128(10,0) literal 1 rescaled to scale 2 */
  i128* f_qty = malloc(BATCH * sizeof(i128));
  i128* f_price = malloc(BATCH * sizeof(i128));
  i128* f_disc = malloc(BATCH * sizeof(i128));
  i128* f_tax = malloc(BATCH * sizeof(i128));
  i128* ce1 = malloc(BATCH * sizeof(i128));
  i128* charge = malloc(BATCH * sizeof(i128));
  uint8_t* f_rf = malloc(BATCH * 16);
  uint8_t* f_ls = malloc(BATCH * 16);
  int32_t* sel = malloc(BATCH * sizeof(int32_t));
  int32_t* gid = malloc(BATCH * sizeof(int32_t));
  p->n_groups = 0;
  for (int64_t b0 = p->row0; b0 < p->row1; b0 += BATCH) {
    const int64_t n = (p->row1 - b0) < BATCH ? (p->row1 - b0) : BATCH;
    /* FilterExec: predicate -> selection vector */
    int m = 0;
    for (int64_t i = 0; i < n; ++i)
      if (p->shipdate[b0 + i] <= p->cutoff) sel[m++] = (int32_t)i;
    /* filter kernel: gather each projected column */
    for (int i = 0; i < m; ++i) f_qty[i] = p->qty[b0 + sel[i]];
    for (int i = 0; i < m; ++i) f_price[i] = p->price[b0 + sel[i]];
    for (int i = 0; i < m; ++i) f_disc[i] = p->disc[b0 + sel[i]];
    for (int i = 0; i < m; ++i) f_tax[i] = p->tax[b0 + sel[i]];
    for (int i = 0; i < m; ++i) memcpy(f_rf + 16 * i, p->rflag + 16 * (b0 + sel[i]), 16);
    for (int i = 0; i < m; ++i) memcpy(f_ls + 16 * i, p->lstatus + 16 * (b0 + sel[i]), 16);
    /* ProjectionExec: __common_expr_1 = price * (1 - disc)  -> Decimal128(32,4) */
    for (int i = 0; i < m; ++i) ce1[i] = f_price[i] * (ONE - f_disc[i]);
    /* aggregate argument: __common_expr_1 * (1 + tax) -> Decimal128(38,6) */
    for (int i = 0; i < m; ++i) charge[i] = ce1[i] * (ONE + f_tax[i]);
    /* AggregateExec(Partial): intern group keys, then one accumulate loop per aggregate */
    for (int i = 0; i < m; ++i) gid[i] = find_group(p, f_rf + 16 * i, f_ls + 16 * i);
    for (int i = 0; i < m; ++i) p->sum_qty[gid[i]] += f_qty[i];
    for (int i = 0; i < m; ++i) p->sum_price[gid[i]] += f_price[i];
    for (int i = 0; i < m; ++i) p->sum_disc_price[gid[i]] += ce1[i];
    for (int i = 0; i < m; ++i) p->sum_charge[gid[i]] += charge[i];
    for (int i = 0; i < m; ++i) p->sum_disc[gid[i]] += f_disc[i];
    for (int i = 0; i < m; ++i) p->count[gid[i]] += 1;
  }
  free(f_qty); free(f_price); free(f_disc); free(f_tax); free(ce1); free(charge); free(f_rf); free(f_ls); free(sel); free(gid);
  return NULL;
}

/*
 * Q1 over rows [0, n).  Outputs (caller arrays of MAXG): key views, 5 sums (i128), counts.
 * Returns the number of groups.  `threads` partitions run concurrently (contiguous row ranges).
 */
int q1_cpu(int64_t n, const void* qty, const void* price, const void* disc, const void* tax, const void* rflag,
           const void* lstatus, const int32_t* shipdate, int32_t cutoff, int threads, uint8_t* out_keys, i128* out_sums,
           int64_t* out_counts) {
  if (threads < 1) threads = 1;
  q1_part* parts = calloc((size_t)threads, sizeof(q1_part));
  pthread_t* tids = malloc((size_t)threads * sizeof(pthread_t));
  const int64_t per = ((n + threads - 1) / threads + BATCH - 1) / BATCH * BATCH;
  for (int t = 0; t < threads; ++t) {
    q1_part* p = &parts[t];
    p->qty = qty; p->price = price; p->disc = disc; p->tax = tax; p->rflag = rflag; p->lstatus = lstatus; p->shipdate = shipdate;
    p->cutoff = cutoff;
    p->row0 = per * t < n ? per * t : n;
    p->row1 = per * (t + 1) < n ? per * (t + 1) : n;
    pthread_create(&tids[t], NULL, q1_task, p);
  }
  for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
  /* FinalPartitioned: merge partial states by key */
  q1_part fin;
  fin.n_groups = 0;
  for (int t = 0; t < threads; ++t) {
    q1_part* p = &parts[t];
    for (int g = 0; g < p->n_groups; ++g) {
      int f = find_group(&fin, p->key[g], p->key[g] + 16);
      fin.sum_qty[f] += p->sum_qty[g]; fin.sum_price[f] += p->sum_price[g]; fin.sum_disc_price[f] += p->sum_disc_price[g];
      fin.sum_charge[f] += p->sum_charge[g]; fin.sum_disc[f] += p->sum_disc[g]; fin.count[f] += p->count[g];
    }
  }
  for (int g = 0; g < fin.n_groups; ++g) {
    memcpy(out_keys + 32 * g, fin.key[g], 32);
    out_sums[5 * g + 0] = fin.sum_qty[g]; out_sums[5 * g + 1] = fin.sum_price[g]; out_sums[5 * g + 2] = fin.sum_disc_price[g];
    out_sums[5 * g + 3] = fin.sum_charge[g]; out_sums[5 * g + 4] = fin.sum_disc[g];
    out_counts[g] = fin.count[g];
  }
  free(parts); free(tids);
  return fin.n_groups;
}

/* ---- Q6: FilterExec(5 conjuncts) -> ProjectionExec(price * disc) -> AggregateExec(no keys) ---- */
typedef struct {
  const i128 *qty, *price, *disc;
  const int32_t* shipdate;
  int64_t row0, row1;
  int32_t d0, d1;
  i128 sum;
  int64_t rows;
} q6_part;

static void* q6_task(void* arg) {
  q6_part* p = (q6_part*)arg;
  int32_t* sel = malloc(BATCH * sizeof(int32_t));
  i128* f_price = malloc(BATCH * sizeof(i128));
  i128* f_disc = malloc(BATCH * sizeof(i128));
  i128* rev = malloc(BATCH * sizeof(i128));
  p->sum = 0; p->rows = 0;
  for (int64_t b0 = p->row0; b0 < p->row1; b0 += BATCH) {
    const int64_t n = (p->row1 - b0) < BATCH ? (p->row1 - b0) : BATCH;
    int m = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t r = b0 + i;
      if (p->shipdate[r] >= p->d0 && p->shipdate[r] < p->d1 && p->disc[r] >= 3 && p->disc[r] <= 5 && p->qty[r] < 2400) sel[m++] = (int32_t)i;
    }
    for (int i = 0; i < m; ++i) f_price[i] = p->price[b0 + sel[i]];
    for (int i = 0; i < m; ++i) f_disc[i] = p->disc[b0 + sel[i]];
    for (int i = 0; i < m; ++i) rev[i] = f_price[i] * f_disc[i];
    for (int i = 0; i < m; ++i) p->sum += rev[i];
    p->rows += m;
  }
  free(sel); free(f_price); free(f_disc); free(rev);
  return NULL;
}

/* returns matching row count; *out_sum = sum(price*disc) (scale 4); 0 rows => SQL NULL */
int64_t q6_cpu(int64_t n, const void* qty, const void* price, const void* disc, const int32_t* shipdate, int32_t d0, int32_t d1,
               int threads, i128* out_sum) {
  if (threads < 1) threads = 1;
  q6_part* parts = calloc((size_t)threads, sizeof(q6_part));
  pthread_t* tids = malloc((size_t)threads * sizeof(pthread_t));
  const int64_t per = ((n + threads - 1) / threads + BATCH - 1) / BATCH * BATCH;
  for (int t = 0; t < threads; ++t) {
    q6_part* p = &parts[t];
    p->qty = qty; p->price = price; p->disc = disc; p->shipdate = shipdate; p->d0 = d0; p->d1 = d1;
    p->row0 = per * t < n ? per * t : n;
    p->row1 = per * (t + 1) < n ? per * (t + 1) : n;
    pthread_create(&tids[t], NULL, q6_task, p);
  }
  i128 sum = 0; int64_t rows = 0;
  for (int t = 0; t < threads; ++t) { pthread_join(tids[t], NULL); sum += parts[t].sum; rows += parts[t].rows; }
  *out_sum = sum;
  free(parts); free(tids);
  return rows;
}
