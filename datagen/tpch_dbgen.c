/*
 * datagen/tpch_dbgen.c -- synthetic-data generator (TPC-H dbgen restatement) used by tests and bench.py.
 * It is neither the product path nor the oracle: it only fabricates input tables.
 *
 * A from-scratch restatement of the TPC-H dbgen (v3) algorithm for the columns the hot path
 * consumes.  The reference's golden vectors (python/pysail/tests/spark/test_tpch.py:11-26,
 * `CALL dbgen(sf = 0.001)` through DuckDB) are produced by TPC-H dbgen; dbgen itself is NOT in
 * /root/reference (it is a DuckDB extension, third party), so the published TPC-H algorithm is
 * restated here: Park-Miller "minimal standard" generator (a = 16807, m = 2^31-1), one seeded
 * stream per column, fixed per-row stream advance ("boundary"), sparse order keys, the
 * PART_SUPP_BRIDGE supplier formula and the retail-price formula of TPC-H spec 4.2.3.
 *
 * Pinned by tests/test_oracle_golden.py: the generated SF0.001 tables fed through the oracle
 * operators reproduce the reference's snapshots test_tpch.result.yaml (Q1,Q3,Q4,Q5,Q6,Q12...).
 *
 * Text columns (comments, addresses, names) are NOT generated: no query on the covered path
 * reads them.  Every stream is independent, so skipping them does not perturb the others.
 *
 * Plain C, no dependencies.  All outputs go to caller-provided arrays.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MODULUS 2147483647LL
#define MULT 16807LL

/* dbgen stream ids (rnd.h Seed[] order) and their initial seeds / per-row boundaries */
enum {
    P_MFG_SD = 0, P_BRND_SD, P_TYPE_SD, P_SIZE_SD, P_CNTR_SD, P_RCST_SD, P_CMNT_SD,
    PS_QTY_SD, PS_SCST_SD, PS_CMNT_SD, O_SUPP_SD, O_CLRK_SD, O_CMNT_SD, O_ODATE_SD,
    L_QTY_SD, L_DCNT_SD, L_TAX_SD, L_SHIP_SD, L_SMODE_SD, L_PKEY_SD, L_SKEY_SD,
    L_SDTE_SD, L_CDTE_SD, L_RDTE_SD, L_RFLG_SD, L_CMNT_SD, C_ADDR_SD, C_NTRG_SD,
    C_PHNE_SD, C_ABAL_SD, C_MSEG_SD, C_CMNT_SD, S_ADDR_SD, S_NTRG_SD, S_PHNE_SD,
    S_ABAL_SD, S_CMNT_SD, P_NAME_SD, O_PRIO_SD, HVAR_SD, O_CKEY_SD, N_CMNT_SD,
    R_CMNT_SD, O_LCNT_SD, BBB_OFFSET_SD, BBB_TYPE_SD, BBB_CMNT_SD, BBB_JNK_SD, N_STREAMS
};

static const int64_t SEED0[N_STREAMS] = {
    1, 46831694, 1841581359, 1193163244, 727633698, 933588178, 804159733,
    1671059989, 1051288424, 1961692154, 1227283347, 1171034773, 276090261, 1066728069,
    209208115, 554590007, 721958466, 1371272478, 675466456, 1808217256, 2095021727,
    1769349045, 904914315, 373135028, 717419739, 1095462486, 881155353, 1489529863,
    1521138112, 298370230, 1140279430, 1335826707, 706178559, 110356601, 884434366,
    962338209, 1341315363, 709314158, 591449447, 431918286, 851767375, 606179079,
    1500869201, 1434868289, 263032577, 753643799, 202794285, 715851524
};

/* values consumed from a stream per generated row (dbgen "boundary") */
static const int BOUNDARY[N_STREAMS] = {
    1, 1, 1, 1, 1, 1, 2,
    4, 4, 8, 1, 1, 2, 1,
    7, 7, 7, 7, 7, 7, 7,
    7, 7, 7, 7, 14, 9, 1,
    3, 1, 1, 2, 9, 1, 3,
    1, 2, 92, 1, 1, 1, 2,
    2, 1, 1, 1, 2, 1
};

static inline int64_t next_rand(int64_t s) { return (s * MULT) % MODULUS; }

/* seed * 16807^n mod m : dbgen's NthElement skip-ahead */
static int64_t nth_element(int64_t n, int64_t seed) {
    int64_t mult = MULT, z = seed;
    while (n > 0) {
        if (n & 1) z = (mult * z) % MODULUS;
        n >>= 1;
        mult = (mult * mult) % MODULUS;
    }
    return z;
}

typedef struct { int64_t v; int used; int boundary; } stream_t;

static void stream_init(stream_t *s, int id, int64_t row0) {
    s->boundary = BOUNDARY[id];
    s->v = nth_element(row0 * (int64_t)BOUNDARY[id], SEED0[id]);
    s->used = 0;
}
/* dbgen UnifInt: low + (long)(seed / m * range), double arithmetic */
static inline int64_t stream_uniform(stream_t *s, int64_t lo, int64_t hi) {
    s->v = next_rand(s->v);
    s->used++;
    double range = (double)(hi - lo + 1);
    return lo + (int64_t)(((double)s->v / 2147483647.0) * range);
}
/* dbgen row_stop: pad the stream to its per-row boundary */
static inline void stream_row_stop(stream_t *s) {
    if (s->used < s->boundary) s->v = nth_element(s->boundary - s->used, s->v);
    s->used = 0;
}

/* ---- scale handling (dbgen main(): bases scaled for sf < 1, integer scale otherwise) ---- */
typedef struct { int64_t part, supp, cust, orders; } tpch_counts;

static tpch_counts counts_for(double sf) {
    tpch_counts c;
    if (sf < 1.0) {
        c.part = (int64_t)(200000 * sf); c.supp = (int64_t)(10000 * sf);
        c.cust = (int64_t)(150000 * sf); c.orders = (int64_t)(1500000 * sf);
        /* dbgen scales tdefs[ORDER].base=150000 then multiplies by ORDERS_PER_CUST */
        c.orders = ((int64_t)(150000 * sf)) * 10;
        if (c.part < 1) c.part = 1;
        if (c.supp < 1) c.supp = 1;
        if (c.cust < 1) c.cust = 1;
        if (c.orders < 1) c.orders = 1;
    } else {
        int64_t s = (int64_t)sf;
        c.part = 200000 * s; c.supp = 10000 * s; c.cust = 150000 * s; c.orders = 1500000 * s;
    }
    return c;
}

void tpch_counts_get(double sf, int64_t *out4) {
    tpch_counts c = counts_for(sf);
    out4[0] = c.part; out4[1] = c.supp; out4[2] = c.cust; out4[3] = c.orders;
}

/* ---- calendar: day index 0 == 1992-01-01 (dbgen STARTDATE); Arrow Date32 = days since 1970 ---- */
#define DATE32_1992_01_01 8035
#define O_ODATE_SPAN (2557 - 151 - 1)   /* TOTDATE - (L_SDTE_MAX + L_RDTE_MAX) - 1 */
#define CURRENT_IDX 1263                /* 1995-06-17 (dbgen CURRENTDATE 95168) */

static inline int64_t sparse_key(int64_t i) {   /* mk_sparse, seq = 0 */
    return ((i >> 3) << 5) | (i & 7);
}
static inline int64_t retail_price(int64_t p) { /* rpb_routine, cents */
    return 90000 + ((p / 10) % 20001) + (p % 1000) * 100;
}

void tpch_retail_price(const int64_t *partkey, int64_t n, int64_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = retail_price(partkey[i]);
}

/*
 * orders + lineitem for order rows [first, first+n) (0-based row index; dbgen index = row+1).
 * Any output pointer may be NULL.  Lineitem arrays must hold 7*n rows; returns #lineitems.
 * Money values are int64 in cents (scale 2); dates are Arrow Date32; flags are single bytes;
 * categorical columns are dictionary codes (see tpch_dict()).
 */
int64_t tpch_gen_orders(double sf, int64_t first, int64_t n,
                        int64_t *o_orderkey, int64_t *o_custkey, uint8_t *o_orderstatus,
                        int64_t *o_totalprice, int32_t *o_orderdate, uint8_t *o_orderpriority,
                        int64_t *o_clerk, int32_t *o_shippriority,
                        int64_t *l_orderkey, int64_t *l_partkey, int64_t *l_suppkey,
                        int32_t *l_linenumber, int64_t *l_quantity, int64_t *l_extendedprice,
                        int64_t *l_discount, int64_t *l_tax, uint8_t *l_returnflag,
                        uint8_t *l_linestatus, int32_t *l_shipdate, int32_t *l_commitdate,
                        int32_t *l_receiptdate, uint8_t *l_shipinstruct, uint8_t *l_shipmode) {
    tpch_counts c = counts_for(sf);
    int64_t scale = sf < 1.0 ? 1 : (int64_t)sf;
    stream_t ckey, odate, prio, clrk, lcnt, qty, dcnt, tax, ship, smode, pkey, skey, sdte, cdte, rdte, rflg;
    stream_init(&ckey, O_CKEY_SD, first); stream_init(&odate, O_ODATE_SD, first);
    stream_init(&prio, O_PRIO_SD, first); stream_init(&clrk, O_CLRK_SD, first);
    stream_init(&lcnt, O_LCNT_SD, first); stream_init(&qty, L_QTY_SD, first);
    stream_init(&dcnt, L_DCNT_SD, first); stream_init(&tax, L_TAX_SD, first);
    stream_init(&ship, L_SHIP_SD, first); stream_init(&smode, L_SMODE_SD, first);
    stream_init(&pkey, L_PKEY_SD, first); stream_init(&skey, L_SKEY_SD, first);
    stream_init(&sdte, L_SDTE_SD, first); stream_init(&cdte, L_CDTE_SD, first);
    stream_init(&rdte, L_RDTE_SD, first); stream_init(&rflg, L_RFLG_SD, first);
    int64_t nl = 0;
    int64_t clerk_max = scale * 1000; if (clerk_max < 1000) clerk_max = 1000;
    for (int64_t r = 0; r < n; r++) {
        int64_t idx = first + r + 1;
        int64_t okey = sparse_key(idx);
        int64_t ck = stream_uniform(&ckey, 1, c.cust);
        int delta = 1;
        while (ck % 3 == 0) { ck += delta; if (ck > c.cust) ck = c.cust; delta = -delta; }
        int64_t od = stream_uniform(&odate, 0, O_ODATE_SPAN - 0);
        int64_t pr = stream_uniform(&prio, 1, 5) - 1;
        int64_t ck_num = stream_uniform(&clrk, 1, clerk_max);
        int64_t lines = stream_uniform(&lcnt, 1, 7);
        int64_t total = 0; int ocnt = 0;
        for (int64_t l = 0; l < lines; l++) {
            int64_t q = stream_uniform(&qty, 1, 50);
            int64_t d = stream_uniform(&dcnt, 0, 10);
            int64_t t = stream_uniform(&tax, 0, 8);
            int64_t si = stream_uniform(&ship, 1, 4) - 1;
            int64_t sm = stream_uniform(&smode, 1, 7) - 1;
            int64_t pk = stream_uniform(&pkey, 1, c.part);
            int64_t rp = retail_price(pk);
            int64_t sn = stream_uniform(&skey, 0, 3);
            int64_t sk = (pk + sn * (c.supp / 4 + (pk - 1) / c.supp)) % c.supp + 1;
            int64_t ep = rp * q;
            total += ((ep * (100 - d)) / 100) * (100 + t) / 100;
            int64_t sd = stream_uniform(&sdte, 1, 121) + od;
            int64_t cd = stream_uniform(&cdte, 30, 90) + od;
            int64_t rd = stream_uniform(&rdte, 1, 30) + sd;
            uint8_t rf = 'N';
            if (rd <= CURRENT_IDX) rf = (stream_uniform(&rflg, 1, 2) == 1) ? 'R' : 'A';
            uint8_t ls = 'O';
            if (sd <= CURRENT_IDX) { ocnt++; ls = 'F'; }
            if (l_orderkey) l_orderkey[nl] = okey;
            if (l_partkey) l_partkey[nl] = pk;
            if (l_suppkey) l_suppkey[nl] = sk;
            if (l_linenumber) l_linenumber[nl] = (int32_t)(l + 1);
            if (l_quantity) l_quantity[nl] = q * 100;
            if (l_extendedprice) l_extendedprice[nl] = ep;
            if (l_discount) l_discount[nl] = d;
            if (l_tax) l_tax[nl] = t;
            if (l_returnflag) l_returnflag[nl] = rf;
            if (l_linestatus) l_linestatus[nl] = ls;
            if (l_shipdate) l_shipdate[nl] = (int32_t)(DATE32_1992_01_01 + sd);
            if (l_commitdate) l_commitdate[nl] = (int32_t)(DATE32_1992_01_01 + cd);
            if (l_receiptdate) l_receiptdate[nl] = (int32_t)(DATE32_1992_01_01 + rd);
            if (l_shipinstruct) l_shipinstruct[nl] = (uint8_t)si;
            if (l_shipmode) l_shipmode[nl] = (uint8_t)sm;
            nl++;
        }
        if (o_orderkey) o_orderkey[r] = okey;
        if (o_custkey) o_custkey[r] = ck;
        if (o_orderstatus) o_orderstatus[r] = ocnt == 0 ? 'O' : (ocnt == lines ? 'F' : 'P');
        if (o_totalprice) o_totalprice[r] = total;
        if (o_orderdate) o_orderdate[r] = (int32_t)(DATE32_1992_01_01 + od);
        if (o_orderpriority) o_orderpriority[r] = (uint8_t)pr;
        if (o_clerk) o_clerk[r] = ck_num;
        if (o_shippriority) o_shippriority[r] = 0;
        stream_row_stop(&ckey); stream_row_stop(&odate); stream_row_stop(&prio);
        stream_row_stop(&clrk); stream_row_stop(&lcnt); stream_row_stop(&qty);
        stream_row_stop(&dcnt); stream_row_stop(&tax); stream_row_stop(&ship);
        stream_row_stop(&smode); stream_row_stop(&pkey); stream_row_stop(&skey);
        stream_row_stop(&sdte); stream_row_stop(&cdte); stream_row_stop(&rdte);
        stream_row_stop(&rflg);
    }
    return nl;
}

/* customer rows [first, first+n): c_custkey, c_nationkey, c_acctbal (cents), c_mktsegment code */
void tpch_gen_customer(double sf, int64_t first, int64_t n, int64_t *c_custkey,
                       int64_t *c_nationkey, int64_t *c_acctbal, uint8_t *c_mktsegment) {
    (void)sf;
    stream_t ntrg, abal, mseg;
    stream_init(&ntrg, C_NTRG_SD, first); stream_init(&abal, C_ABAL_SD, first);
    stream_init(&mseg, C_MSEG_SD, first);
    for (int64_t r = 0; r < n; r++) {
        if (c_custkey) c_custkey[r] = first + r + 1;
        int64_t nk = stream_uniform(&ntrg, 0, 24);
        int64_t ab = stream_uniform(&abal, -99999, 999999);
        int64_t ms = stream_uniform(&mseg, 1, 5) - 1;
        if (c_nationkey) c_nationkey[r] = nk;
        if (c_acctbal) c_acctbal[r] = ab;
        if (c_mktsegment) c_mktsegment[r] = (uint8_t)ms;
        stream_row_stop(&ntrg); stream_row_stop(&abal); stream_row_stop(&mseg);
    }
}

/* supplier rows: s_suppkey, s_nationkey, s_acctbal (cents) */
void tpch_gen_supplier(double sf, int64_t first, int64_t n, int64_t *s_suppkey,
                       int64_t *s_nationkey, int64_t *s_acctbal) {
    (void)sf;
    stream_t ntrg, abal;
    stream_init(&ntrg, S_NTRG_SD, first); stream_init(&abal, S_ABAL_SD, first);
    for (int64_t r = 0; r < n; r++) {
        if (s_suppkey) s_suppkey[r] = first + r + 1;
        int64_t nk = stream_uniform(&ntrg, 0, 24);
        int64_t ab = stream_uniform(&abal, -99999, 999999);
        if (s_nationkey) s_nationkey[r] = nk;
        if (s_acctbal) s_acctbal[r] = ab;
        stream_row_stop(&ntrg); stream_row_stop(&abal);
    }
}

/* part rows: p_partkey, p_retailprice (cents), p_size, p_type code (0..149), p_brand (MN), p_container code */
void tpch_gen_part(double sf, int64_t first, int64_t n, int64_t *p_partkey,
                   int64_t *p_retailprice, int32_t *p_size, uint8_t *p_type,
                   int32_t *p_brand, uint8_t *p_container) {
    (void)sf;
    stream_t mfg, brnd, type, size, cntr;
    stream_init(&mfg, P_MFG_SD, first); stream_init(&brnd, P_BRND_SD, first);
    stream_init(&type, P_TYPE_SD, first); stream_init(&size, P_SIZE_SD, first);
    stream_init(&cntr, P_CNTR_SD, first);
    for (int64_t r = 0; r < n; r++) {
        int64_t pk = first + r + 1;
        int64_t m = stream_uniform(&mfg, 1, 5);
        int64_t b = stream_uniform(&brnd, 1, 5);
        int64_t ty = stream_uniform(&type, 1, 150) - 1;
        int64_t sz = stream_uniform(&size, 1, 50);
        int64_t ct = stream_uniform(&cntr, 1, 40) - 1;
        if (p_partkey) p_partkey[r] = pk;
        if (p_retailprice) p_retailprice[r] = retail_price(pk);
        if (p_size) p_size[r] = (int32_t)sz;
        if (p_type) p_type[r] = (uint8_t)ty;
        if (p_brand) p_brand[r] = (int32_t)(m * 10 + b);
        if (p_container) p_container[r] = (uint8_t)ct;
        stream_row_stop(&mfg); stream_row_stop(&brnd); stream_row_stop(&type);
        stream_row_stop(&size); stream_row_stop(&cntr);
    }
}

/* partsupp rows of parts [first, first+n): 4 suppliers per part (SUPP_PER_PART); arrays hold 4*n rows.
 * ps_suppkey: PART_SUPP_BRIDGE (the formula lineitem uses for l_suppkey); ps_availqty = U(1, 9999) on PS_QTY_SD;
 * ps_supplycost = U(100, 100000) cents on PS_SCST_SD; both streams advance 4 per part row. */
void tpch_gen_partsupp(double sf, int64_t first, int64_t n, int64_t *ps_partkey, int64_t *ps_suppkey,
                       int32_t *ps_availqty, int64_t *ps_supplycost) {
    tpch_counts c = counts_for(sf);
    stream_t qty, scst;
    stream_init(&qty, PS_QTY_SD, first); stream_init(&scst, PS_SCST_SD, first);
    for (int64_t r = 0; r < n; r++) {
        int64_t pk = first + r + 1;
        for (int64_t s = 0; s < 4; s++) {
            int64_t q = stream_uniform(&qty, 1, 9999);
            int64_t cst = stream_uniform(&scst, 100, 100000);
            int64_t k = 4 * r + s;
            if (ps_partkey) ps_partkey[k] = pk;
            if (ps_suppkey) ps_suppkey[k] = (pk + s * (c.supp / 4 + (pk - 1) / c.supp)) % c.supp + 1;
            if (ps_availqty) ps_availqty[k] = (int32_t)q;
            if (ps_supplycost) ps_supplycost[k] = cst;
        }
        stream_row_stop(&qty); stream_row_stop(&scst);
    }
}

/* phone numbers (gen_phone): "CC-AAA-EEE-NNNN", CC = 10 + nation key, then three uniform draws from the phone stream
 * (3 per row).  which = 0: customer (C_PHNE_SD), 1: supplier (S_PHNE_SD).  out: 15 bytes per row, no terminator. */
void tpch_gen_phone(int which, int64_t first, int64_t n, const int64_t *nationkey, char *out) {
    stream_t ph;
    stream_init(&ph, which == 0 ? C_PHNE_SD : S_PHNE_SD, first);
    for (int64_t r = 0; r < n; r++) {
        int64_t ac = stream_uniform(&ph, 100, 999);
        int64_t ex = stream_uniform(&ph, 100, 999);
        int64_t nu = stream_uniform(&ph, 1000, 9999);
        char buf[32];
        int cc = (int)(10 + nationkey[r]);
        buf[0] = (char)('0' + cc / 10); buf[1] = (char)('0' + cc % 10); buf[2] = '-';
        buf[3] = (char)('0' + ac / 100); buf[4] = (char)('0' + ac / 10 % 10); buf[5] = (char)('0' + ac % 10); buf[6] = '-';
        buf[7] = (char)('0' + ex / 100); buf[8] = (char)('0' + ex / 10 % 10); buf[9] = (char)('0' + ex % 10); buf[10] = '-';
        buf[11] = (char)('0' + nu / 1000); buf[12] = (char)('0' + nu / 100 % 10); buf[13] = (char)('0' + nu / 10 % 10); buf[14] = (char)('0' + nu % 10);
        memcpy(out + 15 * r, buf, 15);
        stream_row_stop(&ph);
    }
}

/* p_name (mk_part: agg_str(&colors, 5, P_NAME_SD)): the 92 colour words are permuted per row with a Fisher-Yates pass over the whole
 * list (92 draws: the stream's boundary) and the first five are joined by blanks.  out5: 5 colour indices per row.
 * reset = 1 starts every row from the identity permutation (seedless parallel generation; what the golden results need). */
void tpch_gen_pname(int64_t first, int64_t n, int reset, uint8_t *out5) {
    stream_t nm;
    stream_init(&nm, P_NAME_SD, first);
    int perm[92];
    for (int i = 0; i < 92; i++) perm[i] = i;
    for (int64_t r = 0; r < n; r++) {
        if (reset) for (int i = 0; i < 92; i++) perm[i] = i;
        for (int i = 0; i < 92; i++) {
            int64_t src = stream_uniform(&nm, i, 91);
            int t = perm[src]; perm[src] = perm[i]; perm[i] = t;
        }
        for (int i = 0; i < 5; i++) out5[5 * r + i] = (uint8_t)perm[i];
        stream_row_stop(&nm);
    }
}

/* c_address / s_address (V_STR(25) = a_rnd(10, 40)): one length draw, then one draw per five characters, six bits per character
 * from dbgen's alpha_num table.  which = 0: customer (C_ADDR_SD), 1: supplier (S_ADDR_SD).  out40: 40 bytes per row. */
void tpch_gen_address(int which, int64_t first, int64_t n, int32_t *len, char *out40) {
    static const char alpha_num[] = "0123456789abcdefghijklmnopqrstuvwxyz ABCDEFGHIJKLMNOPQRSTUVWXYZ,";
    stream_t ad;
    stream_init(&ad, which == 0 ? C_ADDR_SD : S_ADDR_SD, first);
    for (int64_t r = 0; r < n; r++) {
        int64_t l = stream_uniform(&ad, 10, 40), bits = 0;
        for (int64_t i = 0; i < l; i++) {
            if (i % 5 == 0) bits = stream_uniform(&ad, 0, 2147483647);
            out40[40 * r + i] = alpha_num[bits & 077];
            bits >>= 6;
        }
        len[r] = (int32_t)l;
        stream_row_stop(&ad);
    }
}

/* supplier "Better Business Bureau" marks (mk_supp): bad_press = U(1, 10000) on BBB_CMNT_SD; at most 10 in 10000 suppliers carry
 * "Customer ... Complaints" (type draw U(0,100) on BBB_TYPE_SD below 50) or "Customer ... Recommends".  kind: 0 none, 1, 2. */
void tpch_gen_bbb(int64_t first, int64_t n, uint8_t *kind) {
    stream_t bp, ty;
    stream_init(&bp, BBB_CMNT_SD, first); stream_init(&ty, BBB_TYPE_SD, first);
    for (int64_t r = 0; r < n; r++) {
        int64_t bad = stream_uniform(&bp, 1, 10000), t = stream_uniform(&ty, 0, 100);
        kind[r] = bad <= 10 ? (t < 50 ? 1 : 2) : 0;
        stream_row_stop(&bp); stream_row_stop(&ty);
    }
}

/* comment columns (TEXT(avg) = dbg_text): dbgen cuts every comment out of one pre-generated text pool -- offset = U(0, pool - max),
 * length = U(min, max), both on the column's stream (two draws per row).  The pool itself needs dbgen's grammar tables, which this
 * repository does not restate: the draws are dbgen's, the pool is datagen/tpch.py's own (see text_pool()).  stream: the Seed[] index. */
void tpch_gen_text(int stream, int64_t first, int64_t n, int64_t min, int64_t max, int64_t pool, int64_t *off, int32_t *len) {
    stream_t tx;
    stream_init(&tx, stream, first);
    for (int64_t r = 0; r < n; r++) {
        off[r] = stream_uniform(&tx, 0, pool - max);
        len[r] = (int32_t)stream_uniform(&tx, min, max);
        stream_row_stop(&tx);
    }
}
