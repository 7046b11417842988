/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C99 + pthreads) of the reference's
 * neighbourhood hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the CPU baseline —
 * never as the product.  The product (libpcpx.so) has no CPU fallback and does not link it.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check every function below against
 *   (i)  the reference's own known-answer tests restated in tests/golden/ (SURVEY.md §8c), and
 *   (ii) the UNMODIFIED reference headers compiled into oracle/_ref/libpcp_ref.so
 *        (oracle/ref_bridge.cpp) on seeded random clouds, in this container; the outputs of
 *        that run are committed as fixtures under tests/golden/ so the pin travels.
 * Exception: estimate_normal's eigen-decomposition lives in Eigen 3.3.8 (fetched by the
 * reference's CMakeLists.txt:18-23, absent from /root/reference and from this image).  Its
 * published algorithm (symmetric 3x3 eigen-decomposition, unit eigenvectors, eigenvalues
 * ascending) is restated with a cyclic Jacobi solver in double; it is pinned by the single
 * known-answer the reference holds (test/common/normal_estimation.cpp:12-37) and by
 * numpy.linalg.eigh as an independent second opinion, with the tolerance the north star
 * states (1 - |cos| <= 1e-4).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/include/pcp/).
 *
 * Build: gcc -O2 -ffp-contract=off (NO -march=native: the host has FMA and the reference's
 * baseline x86-64 build evaluates squared_distance without it).
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * scalar primitives
 * ---------------------------------------------------------------------------------------- */

/* common/norm.hpp:102-112  squared_distance(p1, p2): d = p2 - p1, dx*dx + dy*dy + dz*dz
 * evaluated left to right in fp32 (common/norm.hpp:123-141 gives the same bits). */
static inline float sqdist3(const float* p1, const float* p2)
{
    volatile float dx = p2[0] - p1[0];
    volatile float dy = p2[1] - p1[1];
    volatile float dz = p2[2] - p1[2];
    volatile float xx = dx * dx;
    volatile float yy = dy * dy;
    volatile float zz = dz * dz;
    volatile float s  = xx + yy;
    return s + zz;
}

float oracle_squared_distance(const float* p1, const float* p2) { return sqdist3(p1, p2); }

/* common/vector3d_queries.hpp:31-35  |v1 - v2| < eps, strict, fp32 */
static inline int fp_equals(float a, float b, float eps)
{
    float d = fabsf(a - b);
    return d < eps;
}

/* common/vector3d_queries.hpp:48-64  are_vectors_equal */
static inline int vec_equal(const float* a, const float* b, float eps)
{
    return fp_equals(a[0], b[0], eps) && fp_equals(a[1], b[1], eps) && fp_equals(a[2], b[2], eps);
}

int oracle_are_vectors_equal(const float* a, const float* b, float eps)
{
    return vec_equal(a, b, eps);
}

/* common/axis_aligned_bounding_box.hpp:214-251  bounding_box: per-axis strict < / > updates */
void oracle_bounding_box(const float* xyz, size_t n, float* out6)
{
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
        {
            float v = xyz[3 * i + a];
            if (v < mn[a])
                mn[a] = v;
            if (v > mx[a])
                mx[a] = v;
        }
    memcpy(out6, mn, sizeof mn);
    memcpy(out6 + 3, mx, sizeof mx);
}

/* ------------------------------------------------------------------------------------------
 * octree (octree/linked_octree_node.hpp)
 * ---------------------------------------------------------------------------------------- */

typedef struct
{
    float mn[3], mx[3]; /* voxel_grid_ */
    uint32_t depth;     /* max_depth_ of this node (root = params.max_depth) */
    uint32_t n, cap;    /* elements_ */
    uint32_t* elems;
    int32_t child[8]; /* octants_, -1 = null */
} onode_t;

typedef struct oracle_cloud
{
    const float* xyz; /* borrowed copy below */
    float* xyz_own;
    size_t n;
    uint32_t capacity; /* node_capacity */
    onode_t* nodes;
    size_t n_nodes, cap_nodes;
    size_t inserted;
} oracle_cloud;

static int32_t new_node(oracle_cloud* c, const float* mn, const float* mx, uint32_t depth)
{
    if (c->n_nodes == c->cap_nodes)
    {
        c->cap_nodes = c->cap_nodes ? 2 * c->cap_nodes : 1024;
        c->nodes     = (onode_t*)realloc(c->nodes, c->cap_nodes * sizeof(onode_t));
    }
    onode_t* nd = &c->nodes[c->n_nodes];
    memcpy(nd->mn, mn, 12);
    memcpy(nd->mx, mx, 12);
    nd->depth = depth;
    nd->n = nd->cap = 0;
    nd->elems       = NULL;
    for (int i = 0; i < 8; ++i)
        nd->child[i] = -1;
    return (int32_t)c->n_nodes++;
}

static void node_push(onode_t* nd, uint32_t e)
{
    if (nd->n == nd->cap)
    {
        nd->cap   = nd->cap ? 2 * nd->cap : 8;
        nd->elems = (uint32_t*)realloc(nd->elems, nd->cap * sizeof(uint32_t));
    }
    nd->elems[nd->n++] = e;
}

/* common/axis_aligned_bounding_box.hpp:111-124  contains: inclusive on both ends */
static inline int box_contains(const float* mn, const float* mx, const float* p)
{
    return p[0] >= mn[0] && p[1] >= mn[1] && p[2] >= mn[2] && p[0] <= mx[0] && p[1] <= mx[1] &&
           p[2] <= mx[2];
}

/* octree/linked_octree_node.hpp:163-331  insert (written as a loop instead of recursion) */
static int octree_insert(oracle_cloud* c, uint32_t e)
{
    const float* p = c->xyz + 3 * (size_t)e;
    int32_t cur    = 0;
    for (;;)
    {
        onode_t* nd = &c->nodes[cur];
        if (!box_contains(nd->mn, nd->mx, p)) /* :174 */
            return 0;
        if (nd->depth == 1u) /* :184-188 */
        {
            node_push(nd, e);
            return 1;
        }
        if (nd->n < c->capacity) /* :194-199 */
        {
            node_push(nd, e);
            return 1;
        }
        /* common/axis_aligned_bounding_box.hpp:130  center = (min + max) / 2.f */
        float ctr[3];
        for (int a = 0; a < 3; ++a)
            ctr[a] = (nd->mn[a] + nd->mx[a]) / 2.f;
        int o = 0; /* :258-265 strict > */
        if (p[0] > ctr[0])
            o |= 4;
        if (p[1] > ctr[1])
            o |= 2;
        if (p[2] > ctr[2])
            o |= 1;
        if (nd->child[o] < 0)
        {
            float mn[3], mx[3]; /* :320-327 */
            mn[0] = (o & 4) ? ctr[0] : nd->mn[0], mx[0] = (o & 4) ? nd->mx[0] : ctr[0];
            mn[1] = (o & 2) ? ctr[1] : nd->mn[1], mx[1] = (o & 2) ? nd->mx[1] : ctr[1];
            mn[2] = (o & 1) ? ctr[2] : nd->mn[2], mx[2] = (o & 1) ? nd->mx[2] : ctr[2];
            uint32_t d   = nd->depth - 1u; /* :317 */
            int32_t id   = new_node(c, mn, mx, d);
            nd           = &c->nodes[cur]; /* realloc may have moved the array */
            nd->child[o] = id;
        }
        cur = nd->child[o];
    }
}

/* octree/linked_octree.hpp:83-91 (explicit params) and :103-121 (auto bounding box).
 * bbox6 == NULL -> auto bbox.  node_capacity / max_depth 0 -> defaults 32 / 21
 * (octree/linked_octree_node.hpp:37-38). */
oracle_cloud* oracle_cloud_create(
    const float* xyz,
    size_t n,
    const float* bbox6,
    uint32_t node_capacity,
    uint32_t max_depth)
{
    oracle_cloud* c = (oracle_cloud*)calloc(1, sizeof *c);
    c->xyz_own      = (float*)malloc((n ? n : 1) * 12);
    memcpy(c->xyz_own, xyz, n * 12);
    c->xyz      = c->xyz_own;
    c->n        = n;
    c->capacity = node_capacity ? node_capacity : 32u;
    float bb[6];
    if (bbox6)
        memcpy(bb, bbox6, sizeof bb);
    else
        oracle_bounding_box(xyz, n, bb);
    new_node(c, bb, bb + 3, max_depth ? max_depth : 21u);
    for (size_t i = 0; i < n; ++i)
        c->inserted += (size_t)octree_insert(c, (uint32_t)i);
    return c;
}

void oracle_cloud_destroy(oracle_cloud* c)
{
    if (!c)
        return;
    for (size_t i = 0; i < c->n_nodes; ++i)
        free(c->nodes[i].elems);
    free(c->nodes);
    free(c->xyz_own);
    free(c);
}

size_t oracle_cloud_size(const oracle_cloud* c) { return c->inserted; }
size_t oracle_cloud_nodes(const oracle_cloud* c) { return c->n_nodes; }

void oracle_cloud_bbox(const oracle_cloud* c, float* out6)
{
    memcpy(out6, c->nodes[0].mn, 12);
    memcpy(out6 + 3, c->nodes[0].mx, 12);
}

/* ---- best-first kNN: octree/linked_octree_node.hpp:453-570 --------------------------------
 * The reference's heap orders by distance only (:480-488), so the order among bit-equal fp32
 * distances is an accident of std::priority_queue.  The restatement makes that order
 * deterministic — key (d2, nodes-before-points, id) — which yields exactly the tie-aware
 * contract of SURVEY.md §8c: neighbours ascending by (d2, original index). */
typedef struct
{
    float d2;
    uint32_t is_point;
    uint32_t id;
} hent_t;

typedef struct
{
    hent_t* a;
    size_t n, cap;
} heap_t;

static inline int hless(const hent_t* x, const hent_t* y)
{
    if (x->d2 != y->d2)
        return x->d2 < y->d2;
    if (x->is_point != y->is_point)
        return x->is_point < y->is_point;
    return x->id < y->id;
}

static void hpush(heap_t* h, hent_t e)
{
    if (h->n == h->cap)
    {
        h->cap = h->cap ? 2 * h->cap : 256;
        h->a   = (hent_t*)realloc(h->a, h->cap * sizeof(hent_t));
    }
    size_t i = h->n++;
    while (i > 0)
    {
        size_t p = (i - 1) / 2;
        if (!hless(&e, &h->a[p]))
            break;
        h->a[i] = h->a[p];
        i       = p;
    }
    h->a[i] = e;
}

static hent_t hpop(heap_t* h)
{
    hent_t top  = h->a[0];
    hent_t last = h->a[--h->n];
    size_t i    = 0;
    for (;;)
    {
        size_t l = 2 * i + 1, r = l + 1, m;
        if (l >= h->n)
            break;
        m = (r < h->n && hless(&h->a[r], &h->a[l])) ? r : l;
        if (!hless(&h->a[m], &last))
            break;
        h->a[i] = h->a[m];
        i       = m;
    }
    if (h->n)
        h->a[i] = last;
    return top;
}

/* common/axis_aligned_bounding_box.hpp:139-148  nearest_point_from = per-axis clamp */
static inline float box_d2(const onode_t* nd, const float* t)
{
    float q[3];
    for (int a = 0; a < 3; ++a)
    {
        float v = t[a];
        v       = v < nd->mn[a] ? nd->mn[a] : v;
        v       = nd->mx[a] < v ? nd->mx[a] : v;
        q[a]    = v;
    }
    return sqdist3(t, q);
}

static size_t knn_one(
    const oracle_cloud* c,
    heap_t* h,
    const float* t,
    size_t k,
    float eps,
    int64_t* out_idx,
    float* out_d2)
{
    size_t found = 0;
    if (k == 0) /* :464-465 */
        return 0;
    h->n = 0;
    hent_t root = {box_d2(&c->nodes[0], t), 0u, 0u};
    hpush(h, root);
    while (found < k && h->n) /* :525 */
    {
        hent_t e = hpop(h);
        if (e.is_point) /* :536-543 */
        {
            const float* p = c->xyz + 3 * (size_t)e.id;
            if (!vec_equal(p, t, eps))
            {
                out_idx[found] = (int64_t)e.id;
                if (out_d2)
                    out_d2[found] = e.d2;
                ++found;
            }
            continue;
        }
        const onode_t* nd = &c->nodes[e.id];
        for (uint32_t i = 0; i < nd->n; ++i) /* :551-554 */
        {
            hent_t pe = {sqdist3(t, c->xyz + 3 * (size_t)nd->elems[i]), 1u, nd->elems[i]};
            hpush(h, pe);
        }
        for (int o = 0; o < 8; ++o) /* :560-566 */
            if (nd->child[o] >= 0)
            {
                hent_t ne = {box_d2(&c->nodes[nd->child[o]], t), 0u, (uint32_t)nd->child[o]};
                hpush(h, ne);
            }
    }
    return found;
}

/* ---- range search: octree/linked_octree_node.hpp:581-614 ----------------------------------
 * sphere containment common/sphere.hpp:27-35: squared_distance(center, p) <= radius * radius.
 * Pruning common/intersections.hpp:87-102 compares the SQUARED box distance with the
 * UN-squared radius (:101).  exact_prune = 0 restates that literally (for r > 1 the reference
 * can miss in-range points; for r <= 1 the prune is only weaker and results are exact);
 * exact_prune = 1 uses r*r and is the mathematically intended predicate. */
typedef struct
{
    uint32_t* a;
    size_t n, cap;
} u32vec_t;

static void vpush(u32vec_t* v, uint32_t x)
{
    if (v->n == v->cap)
    {
        v->cap = v->cap ? 2 * v->cap : 64;
        v->a   = (uint32_t*)realloc(v->a, v->cap * sizeof(uint32_t));
    }
    v->a[v->n++] = x;
}

static void radius_rec(
    const oracle_cloud* c,
    int32_t ni,
    const float* ctr,
    float r,
    float rr,
    int exact_prune,
    u32vec_t* out)
{
    const onode_t* nd = &c->nodes[ni];
    for (uint32_t i = 0; i < nd->n; ++i) /* :587-589 */
        if (sqdist3(ctr, c->xyz + 3 * (size_t)nd->elems[i]) <= rr)
            vpush(out, nd->elems[i]);
    for (int o = 0; o < 8; ++o) /* :591-613 */
    {
        if (nd->child[o] < 0)
            continue;
        const onode_t* ch = &c->nodes[nd->child[o]];
        int inside        = box_contains(ch->mn, ch->mx, ctr); /* intersections.hpp:92-98 */
        if (!inside)
        {
            /* intersections.hpp:100-101: squared_distance(nearest_point, center) <= s.radius */
            float q[3];
            for (int a = 0; a < 3; ++a)
            {
                float v = ctr[a];
                v       = v < ch->mn[a] ? ch->mn[a] : v;
                v       = ch->mx[a] < v ? ch->mx[a] : v;
                q[a]    = v;
            }
            float d2 = sqdist3(q, ctr);
            if (!(d2 <= (exact_prune ? rr : r)))
                continue;
        }
        radius_rec(c, nd->child[o], ctr, r, rr, exact_prune, out);
    }
}

static int cmp_u32(const void* a, const void* b)
{
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------
 * estimate_normal: common/normals/normal_estimation.hpp:32-78
 * ---------------------------------------------------------------------------------------- */

/* Cyclic Jacobi for a symmetric 3x3 in double.  w ascending, V columns = unit eigenvectors
 * (what Eigen::SelfAdjointEigenSolver publishes: eigenvalues sorted increasing, normalised
 * eigenvectors; Eigen 3.3.8 docs of SelfAdjointEigenSolver::eigenvalues/eigenvectors). */
static void jacobi3(const double A_in[3][3], double w[3], double V[3][3])
{
    double A[3][3];
    memcpy(A, A_in, sizeof A);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            V[i][j] = i == j;
    for (int sweep = 0; sweep < 64; ++sweep)
    {
        double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        double dia = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off == 0.0 || off <= 1e-40 * dia)
            break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q)
            {
                if (A[p][q] == 0.0)
                    continue;
                double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) /* A <- A J */
                {
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = cs * akp - sn * akq;
                    A[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) /* A <- J^T A */
                {
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = cs * apk - sn * aqk;
                    A[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k)
                {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    double d[3] = {A[0][0], A[1][1], A[2][2]};
    for (int i = 0; i < 2; ++i) /* selection sort ascending, as Eigen does */
    {
        int m = i;
        for (int j = i + 1; j < 3; ++j)
            if (d[ord[j]] < d[ord[m]])
                m = j;
        int t = ord[i];
        ord[i] = ord[m];
        ord[m] = t;
    }
    double Vs[3][3];
    for (int j = 0; j < 3; ++j)
    {
        w[j] = d[ord[j]];
        double nrm = 0;
        for (int i = 0; i < 3; ++i)
            nrm += V[i][ord[j]] * V[i][ord[j]];
        nrm = sqrt(nrm);
        for (int i = 0; i < 3; ++i)
            Vs[i][j] = V[i][ord[j]] / nrm;
    }
    memcpy(V, Vs, sizeof Vs);
}

/* The fp32 scatter matrix the reference hands to the eigensolver:
 * :42-48 V (3xn fp32) -> :50 Mu = rowwise mean -> :51 V' = V - Mu -> :52 Cov = V' V'^T
 * (un-normalised).  cov6 = xx, xy, xz, yy, yz, zz; mu3 optional. */
void oracle_scatter_matrix(const float* pts, size_t n, float* cov6, float* mu3)
{
    float mu[3] = {0.f, 0.f, 0.f};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
            mu[a] += pts[3 * i + a];
    for (int a = 0; a < 3; ++a)
        mu[a] = mu[a] / (float)n; /* n == 0 -> NaN, as Eigen's mean() of an empty row */
    float c[6] = {0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i)
    {
        float x = pts[3 * i] - mu[0], y = pts[3 * i + 1] - mu[1], z = pts[3 * i + 2] - mu[2];
        c[0] += x * x, c[1] += x * y, c[2] += x * z;
        c[3] += y * y, c[4] += y * z, c[5] += z * z;
    }
    memcpy(cov6, c, sizeof c);
    if (mu3)
        memcpy(mu3, mu, sizeof mu);
}

/* common/normals/normal_estimation.hpp:41-77.  pts = n neighbour positions (the query is
 * NOT among them because kNN excluded it).  out3 = unit eigenvector of the smallest
 * eigenvalue; on exactly equal eigenvalues the later `if` wins (:60-73).  Sign is not
 * canonical in the reference (whatever Eigen returns); compare with |dot|.
 * out_gap (optional) = (l1 - l0) / max(l2, tiny): small values flag ill-conditioned normals. */
void oracle_estimate_normal(const float* pts, size_t n, float* out3, float* out_gap)
{
    float c[6];
    oracle_scatter_matrix(pts, n, c, NULL);
    double A[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    double w[3], V[3][3];
    jacobi3(A, w, V);
    float l[3] = {(float)w[0], (float)w[1], (float)w[2]}; /* Eigen returns fp32 eigenvalues */
    int col = 0;
    if (l[0] <= l[1] && l[0] <= l[2])
        col = 0;
    if (l[1] <= l[0] && l[1] <= l[2])
        col = 1;
    if (l[2] <= l[0] && l[2] <= l[1])
        col = 2;
    for (int i = 0; i < 3; ++i)
        out3[i] = (float)V[i][col];
    if (out_gap)
        *out_gap = (float)((w[1] - w[0]) / (w[2] > 1e-300 ? w[2] : 1e-300));
}

/* ------------------------------------------------------------------------------------------
 * batched drivers (queries partitioned statically over pthreads: the stand-in for
 * std::execution::par, which is serial in this image — SURVEY.md §8d)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
    const oracle_cloud* c;
    const float* queries;
    size_t q0, q1, k;
    float eps;
    int64_t* idx;
    float* d2;
    uint32_t* count;
    float* normals;
    float* gaps;
    float* means;
    /* radius */
    const float* radii;
    float r;
    int exact_prune;
    const uint64_t* offsets;
    int mode; /* 0 knn, 1 normals, 2 radius, 3 mean distance */
} job_t;

static const float* target_of(const job_t* j, size_t i)
{
    return j->queries ? j->queries + 3 * i : j->c->xyz + 3 * i;
}

static void* job_run(void* arg)
{
    job_t* j       = (job_t*)arg;
    heap_t h       = {0};
    u32vec_t v     = {0};
    size_t k       = j->k;
    int64_t* tmp_i = (int64_t*)malloc((k ? k : 1) * sizeof(int64_t));
    float* tmp_d   = (float*)malloc((k ? k : 1) * sizeof(float));
    float* tmp_p   = (float*)malloc((k ? k : 1) * 12);
    for (size_t i = j->q0; i < j->q1; ++i)
    {
        const float* t = target_of(j, i);
        if (j->mode == 2)
        {
            float r = j->radii ? j->radii[i] : j->r;
            volatile float rr = r * r; /* common/sphere.hpp:34 radius * radius in fp32 */
            v.n = 0;
            radius_rec(j->c, 0, t, r, rr, j->exact_prune, &v);
            if (j->count)
                j->count[i] = (uint32_t)v.n;
            if (j->idx && j->offsets)
            {
                qsort(v.a, v.n, sizeof(uint32_t), cmp_u32);
                for (size_t m = 0; m < v.n; ++m)
                    j->idx[j->offsets[i] + m] = (int64_t)v.a[m];
            }
            continue;
        }
        size_t m = knn_one(j->c, &h, t, k, j->eps, tmp_i, tmp_d);
        if (j->mode == 0)
        {
            for (size_t s = 0; s < k; ++s)
            {
                if (j->idx)
                    j->idx[i * k + s] = s < m ? tmp_i[s] : -1;
                if (j->d2)
                    j->d2[i * k + s] = s < m ? tmp_d[s] : INFINITY;
            }
            if (j->count)
                j->count[i] = (uint32_t)m;
        }
        else if (j->mode == 1)
        {
            /* algorithm/estimate_normals.hpp:80-90: knn(v) -> estimate_normal -> op(v, n) with
             * default_normal_transform (algorithm/common.hpp:31-34) returning n */
            for (size_t s = 0; s < m; ++s)
                memcpy(tmp_p + 3 * s, j->c->xyz + 3 * (size_t)tmp_i[s], 12);
            oracle_estimate_normal(tmp_p, m, j->normals + 3 * i, j->gaps ? j->gaps + i : NULL);
        }
        else
        {
            /* algorithm/average_distance_to_neighbors.hpp:56-70: sequential fp32 sum of
             * norm(pi - pj) (common/norm.hpp:60-70: sqrt(xx + yy + zz)), / neighbours.size() */
            float sum = 0.f;
            for (size_t s = 0; s < m; ++s)
            {
                const float* p = j->c->xyz + 3 * (size_t)tmp_i[s];
                float dx = t[0] - p[0], dy = t[1] - p[1], dz = t[2] - p[2];
                volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                volatile float s2 = xx + yy;
                sum += sqrtf(s2 + zz);
            }
            j->means[i] = sum / (float)m;
        }
    }
    free(h.a);
    free(v.a);
    free(tmp_i);
    free(tmp_d);
    free(tmp_p);
    return NULL;
}

static void run_jobs(job_t proto, size_t nq, int nthreads)
{
    if (nthreads < 1)
        nthreads = 1;
    if ((size_t)nthreads > nq)
        nthreads = nq ? (int)nq : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs   = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t)
    {
        jobs[t]    = proto;
        jobs[t].q0 = nq * (size_t)t / (size_t)nthreads;
        jobs[t].q1 = nq * (size_t)(t + 1) / (size_t)nthreads;
        if (nthreads == 1)
            job_run(&jobs[t]);
        else
            pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t)
            pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* queries == NULL -> the indexed points themselves.  out_idx nq*k (pad -1), out_d2 nq*k
 * (pad +inf) and out_count may each be NULL. */
void oracle_knn(
    const oracle_cloud* c,
    const float* queries,
    size_t nq,
    size_t k,
    double eps,
    int64_t* out_idx,
    float* out_d2,
    uint32_t* out_count,
    int nthreads)
{
    job_t j = {0};
    j.c = c, j.queries = queries, j.k = k, j.eps = (float)eps; /* :540 static_cast<float>(eps) */
    j.idx = out_idx, j.d2 = out_d2, j.count = out_count, j.mode = 0;
    run_jobs(j, nq, nthreads);
}

void oracle_estimate_normals(
    const oracle_cloud* c,
    const float* queries,
    size_t nq,
    size_t k,
    double eps,
    float* out_normals,
    float* out_gaps,
    int nthreads)
{
    job_t j = {0};
    j.c = c, j.queries = queries, j.k = k, j.eps = (float)eps;
    j.normals = out_normals, j.gaps = out_gaps, j.mode = 1;
    run_jobs(j, nq, nthreads);
}

/* two-call protocol for lists: call with out_idx == NULL to get counts, prefix-sum them into
 * offsets (nq + 1), call again with offsets and out_idx; indices ascending per query. */
void oracle_radius(
    const oracle_cloud* c,
    const float* queries,
    size_t nq,
    const float* radii,
    float r,
    int exact_prune,
    uint32_t* out_count,
    const uint64_t* offsets,
    int64_t* out_idx,
    int nthreads)
{
    job_t j = {0};
    j.c = c, j.queries = queries, j.radii = radii, j.r = r, j.exact_prune = exact_prune;
    j.count = out_count, j.offsets = offsets, j.idx = out_idx, j.mode = 2;
    run_jobs(j, nq, nthreads);
}

/* algorithm/average_distance_to_neighbors.hpp:39-113 over the cloud's own points.
 * out_means: n per-point means.  Returns (sum_i mean_i) / n accumulated sequentially in fp32
 * in index order (:108-111; the example's std::reduce(par) order is unspecified,
 * examples/filter_point_cloud_noise_by_density.cpp:73-75). */
float oracle_average_distance_to_neighbors(
    const oracle_cloud* c,
    size_t k,
    double eps,
    float* out_means,
    int nthreads)
{
    float* means = out_means ? out_means : (float*)malloc((c->n ? c->n : 1) * sizeof(float));
    job_t j      = {0};
    j.c = c, j.k = k, j.eps = (float)eps, j.means = means, j.mode = 3;
    run_jobs(j, c->n, nthreads);
    float sum = 0.f;
    for (size_t i = 0; i < c->n; ++i)
        sum += means[i];
    float mu = sum / (float)c->n;
    if (!out_means)
        free(means);
    return mu;
}

/* examples/filter_point_cloud_noise_by_density.cpp:81-91 on the IMMUTABLE input cloud:
 * keep[i] = |range_search(sphere{p_i, radius})| >= threshold; the count includes p_i itself.
 * `radius` is an explicit fp32 input (already multiplied by the example's radius multiplier).
 * Returns the number of kept points. */
size_t oracle_density_filter(
    const oracle_cloud* c,
    float radius,
    uint32_t threshold,
    uint8_t* out_keep,
    uint32_t* out_count,
    int nthreads)
{
    uint32_t* cnt = out_count ? out_count : (uint32_t*)malloc((c->n ? c->n : 1) * 4);
    oracle_radius(c, NULL, c->n, NULL, radius, 0, cnt, NULL, NULL, nthreads);
    size_t kept = 0;
    for (size_t i = 0; i < c->n; ++i)
    {
        int keep = !(cnt[i] < threshold); /* :89 removes when density < threshold */
        if (out_keep)
            out_keep[i] = (uint8_t)keep;
        kept += (size_t)keep;
    }
    if (!out_count)
        free(cnt);
    return kept;
}

/* ------------------------------------------------------------------------------------------
 * brute force (O(n) per query) — an independent second opinion for small cases
 * ---------------------------------------------------------------------------------------- */
void oracle_knn_bruteforce(
    const float* xyz,
    size_t n,
    const float* queries,
    size_t nq,
    size_t k,
    double eps_d,
    int64_t* out_idx,
    float* out_d2,
    uint32_t* out_count)
{
    float eps = (float)eps_d;
    for (size_t i = 0; i < nq; ++i)
    {
        const float* t = queries ? queries + 3 * i : xyz + 3 * i;
        size_t m       = 0;
        for (size_t p = 0; k > 0 && p < n; ++p)
        {
            const float* x = xyz + 3 * p;
            if (vec_equal(x, t, eps))
                continue;
            float d = sqdist3(t, x);
            if (m == k && !(d < out_d2[i * k + k - 1]))
                continue; /* later index never displaces an equal distance */
            size_t s = m < k ? m : k - 1;
            while (s > 0 && out_d2[i * k + s - 1] > d)
            {
                out_d2[i * k + s]  = out_d2[i * k + s - 1];
                out_idx[i * k + s] = out_idx[i * k + s - 1];
                --s;
            }
            out_d2[i * k + s]  = d;
            out_idx[i * k + s] = (int64_t)p;
            if (m < k)
                ++m;
        }
        for (size_t s = m; s < k; ++s)
        {
            out_idx[i * k + s] = -1;
            out_d2[i * k + s]  = INFINITY;
        }
        if (out_count)
            out_count[i] = (uint32_t)m;
    }
}

void oracle_radius_count_bruteforce(
    const float* xyz,
    size_t n,
    const float* queries,
    size_t nq,
    const float* radii,
    float r,
    uint32_t* out_count)
{
    for (size_t i = 0; i < nq; ++i)
    {
        const float* t    = queries ? queries + 3 * i : xyz + 3 * i;
        float ri          = radii ? radii[i] : r;
        volatile float rr = ri * ri;
        uint32_t cnt      = 0;
        for (size_t p = 0; p < n; ++p)
            cnt += sqdist3(t, xyz + 3 * p) <= rr;
        out_count[i] = cnt;
    }
}

/* ------------------------------------------------------------------------------------------
 * Radius-search callers.  The reference gathers each ball with a kd-tree
 * rebuilt per iteration; the restatement gathers it from the octree above (same predicate,
 * sphere_a::contains, common/sphere.hpp:51-56) — the neighbour ORDER therefore differs from
 * the reference's kd-tree DFS order and fp32 sums agree only to rounding.
 *
 * Parity status: bilateral_filter_points and the four WLOP bodies are PINNED against the
 * unmodified reference (oracle/ref_bridge_smoothing.cpp, fixtures in tests/golden/) and the
 * reference's own tests (test/algorithm/bilateral_filter.cpp, test/algorithm/wlop.cpp).
 * bilateral_filter_normals is UNPINNED: compute_ni is written on Eigen types (Eigen 3.3.8,
 * absent here) and the reference's test only checks the output size
 * (test/algorithm/bilateral_filter.cpp:135-155); the formula is restated line by line.
 * ---------------------------------------------------------------------------------------- */

/* algorithm/bilateral_filter.hpp:359-367 */
static float gaussian_f(float sigma, float r)
{
    float const s2    = sigma * sigma;
    float const r2    = r * r;
    float const power = -r2 / (2 * s2);
    float const coeff = 1.f / (sigma * sqrtf(2.f * 3.14159265358979323846f));
    return coeff * expf(power);
}

/* algorithm/bilateral_filter.hpp:511-519 */
static float dgaussian_f(float sigma, float r)
{
    float const s2    = sigma * sigma;
    float const s3    = sigma * s2;
    float const r2    = r * r;
    float const power = -r2 / (2 * s2);
    float const coeff = -r / (s3 * sqrtf(2.f * 3.14159265358979323846f));
    return coeff * expf(power);
}

static u32vec_t ball(const oracle_cloud* c, const float* ctr, float r)
{
    u32vec_t v = {0};
    if (c->n_nodes)
        radius_rec(c, 0, ctr, r, r * r, 1, &v);
    return v;
}

/* bilateral::detail::compute_pi, algorithm/bilateral_filter.hpp:47-100 */
static void bilateral_pi(const oracle_cloud* c, const float* nrm, size_t i, float sigmaf,
                         float sigmag, float* out)
{
    const float* s = c->xyz + 3 * i;
    u32vec_t nb    = ball(c, s, 2.f * sigmaf);
    float k = 0.f, acc[3] = {0.f, 0.f, 0.f};
    for (size_t j = 0; j < nb.n; ++j)
    {
        const float* p = c->xyz + 3 * (size_t)nb.a[j];
        const float* n = nrm + 3 * (size_t)nb.a[j];
        float sp[3]    = {p[0] - s[0], p[1] - s[1], p[2] - s[2]};      /* :372 */
        float d        = sp[0] * n[0] + sp[1] * n[1] + sp[2] * n[2];  /* :374 */
        float pj[3]    = {s[0] + d * n[0], s[1] + d * n[1], s[2] + d * n[2]};
        float f[3]     = {s[0] - p[0], s[1] - p[1], s[2] - p[2]};
        float g[3]     = {pj[0] - s[0], pj[1] - s[1], pj[2] - s[2]};
        float rf       = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]); /* :83 */
        float rg       = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]); /* :84 */
        float w        = gaussian_f(sigmaf, rf) * gaussian_f(sigmag, rg);
        k += w;
        for (int a = 0; a < 3; ++a)
            acc[a] = acc[a] + w * pj[a];
    }
    for (int a = 0; a < 3; ++a)
        out[a] = acc[a] / k;
    free(nb.a);
}

/* algorithm/bilateral_filter.hpp:301-421: K iterations, tree rebuilt on the moved points */
void oracle_bilateral_filter_points(const float* xyz, const float* normals, size_t n, double sigmaf,
                                    double sigmag, size_t K, float* out)
{
    float* cur = (float*)malloc((n ? n : 1) * 12);
    memcpy(cur, xyz, n * 12);
    for (size_t it = 0; it < K; ++it)
    {
        oracle_cloud* c = oracle_cloud_create(cur, n, NULL, 0, 0);
        float* nxt      = (float*)malloc((n ? n : 1) * 12);
        for (size_t i = 0; i < n; ++i)
            bilateral_pi(c, normals, i, (float)sigmaf, (float)sigmag, nxt + 3 * i);
        oracle_cloud_destroy(c);
        free(cur);
        cur = nxt;
    }
    memcpy(out, cur, n * 12);
    free(cur);
}

static void normalized3(const float* v, float* u)
{
    /* Eigen 3.3 MatrixBase::normalized(): v itself unless its squared norm is positive */
    float z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (z > 0.f)
    {
        float s = sqrtf(z);
        u[0] = v[0] / s, u[1] = v[1] / s, u[2] = v[2] / s;
    }
    else
        u[0] = v[0], u[1] = v[1], u[2] = v[2];
}

/* bilateral::detail::compute_ni, algorithm/bilateral_filter.hpp:113-267 */
static void bilateral_ni(const oracle_cloud* c, const float* nrm, size_t i, float sigmaf,
                         float sigmag, float* out)
{
    const float* s = c->xyz + 3 * i;
    u32vec_t nb    = ball(c, s, 2.f * sigmaf);
    float J[3][3] = {{0}}, u[3] = {0}, gk[3] = {0}, k = 0.f;
    for (size_t j = 0; j < nb.n; ++j)
    {
        const float* p = c->xyz + 3 * (size_t)nb.a[j];
        const float* n = nrm + 3 * (size_t)nb.a[j];
        float v[3]     = {p[0] - s[0], p[1] - s[1], p[2] - s[2]};
        float d        = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
        float pj[3]    = {s[0] + d * n[0], s[1] + d * n[1], s[2] + d * n[2]};
        float sp[3]    = {s[0] - p[0], s[1] - p[1], s[2] - p[2]};       /* :184 */
        float sps[3]   = {pj[0] - s[0], pj[1] - s[1], pj[2] - s[2]};    /* :185 */
        float rf       = sqrtf(sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2]);
        float rg       = sqrtf(sps[0] * sps[0] + sps[1] * sps[1] + sps[2] * sps[2]);
        float wf = gaussian_f(sigmaf, rf), wg = gaussian_f(sigmag, rg);
        float w  = wf * wg;
        k += w;
        for (int a = 0; a < 3; ++a)
            u[a] += w * pj[a];
        float wdf = dgaussian_f(sigmaf, rf);
        float su[3], gf[3];
        normalized3(sp, su);
        for (int a = 0; a < 3; ++a)
            gf[a] = su[a] * wdf;
        float P[3][3] = {{1 - n[0] * n[0], n[0] * n[1], n[0] * n[2]},   /* :213-222 */
                         {n[0] * n[1], 1 - n[1] * n[1], n[1] * n[2]},
                         {n[0] * n[2], n[1] * n[2], 1 - n[2] * n[2]}};
        float wdg = dgaussian_f(sigmag, rg);
        float pu[3], gg[3];
        normalized3(sps, pu);
        for (int a = 0; a < 3; ++a)                                      /* :234 */
            gg[a] = ((pu[0] * P[0][a] + pu[1] * P[1][a] + pu[2] * P[2][a]) - pu[a]) * wdg;
        for (int a = 0; a < 3; ++a)                                      /* :239 */
            gk[a] += gf[a] * wg + wf * gg[a];
        for (int r = 0; r < 3; ++r)                                      /* :243 */
            for (int a = 0; a < 3; ++a)
                J[r][a] += (P[r][a] * wf * wg + sps[r] * gf[a] * wg) + sps[r] * wf * gg[a];
    }
    float inv = 1.f / (k * k);                                           /* :249-250 */
    const float* ns = nrm + 3 * i;
    float o[3];
    for (int r = 0; r < 3; ++r)
    {
        float acc = 0.f;
        for (int a = 0; a < 3; ++a)
            acc += (inv * (J[r][a] * k - u[r] * gk[a])) * ns[a];
        o[r] = acc;
    }
    normalized3(o, out);                                                 /* :264-266 */
    free(nb.a);
}

/* algorithm/bilateral_filter.hpp:452-575: one tree, K passes over the normals */
void oracle_bilateral_filter_normals(const float* xyz, const float* normals, size_t n,
                                     double sigmaf, double sigmag, size_t K, float* out)
{
    oracle_cloud* c = oracle_cloud_create(xyz, n, NULL, 0, 0);
    float* cur      = (float*)malloc((n ? n : 1) * 12);
    float* nxt      = (float*)malloc((n ? n : 1) * 12);
    memcpy(cur, normals, n * 12);
    for (size_t it = 0; it < K; ++it)
    {
        for (size_t i = 0; i < n; ++i)
            bilateral_ni(c, cur, i, (float)sigmaf, (float)sigmag, nxt + 3 * i);
        float* t = cur;
        cur = nxt, nxt = t;
    }
    memcpy(out, cur, n * 12);
    free(cur), free(nxt);
    oracle_cloud_destroy(c);
}

/* ---- WLOP: algorithm/wlop.hpp ---- */
static float theta_f(float r2, float h4sq) { return expf(-r2 / h4sq); } /* :326-328 */

static int same_point(const float* a, const float* b) /* are_vectors_equal(a, b, 1e-9f) */
{
    float const eps = (float)1e-9;
    return fp_equals(a[0], b[0], eps) && fp_equals(a[1], b[1], eps) && fp_equals(a[2], b[2], eps);
}

/* compute_vj / compute_wi, :28-104 */
static float wlop_density(const oracle_cloud* c, const float* q, float h, float h4sq)
{
    u32vec_t nb = ball(c, q, h);
    float v     = 1.0f;
    for (size_t j = 0; j < nb.n; ++j)
    {
        const float* p = c->xyz + 3 * (size_t)nb.a[j];
        if (same_point(q, p))
            continue;
        v += theta_f(sqdist3(q, p), h4sq);
    }
    free(nb.a);
    return v;
}

/* initial: I indices into xyz (the reference draws them with std::random_device, :346-358) */
void oracle_wlop(const float* xyz, size_t n, const uint32_t* initial, size_t I, double mu_d,
                 double h_d, size_t K, int uniform, float* out)
{
    float const mu = (float)mu_d, h = (float)h_d;
    float const h4sq = (h * h) / (float)(4.0 * 4.0); /* :322-324 */
    float const eps = (float)1e-9, zero = 0.f;
    float* x  = (float*)malloc((I ? I : 1) * 12);
    float* xp = (float*)malloc((I ? I : 1) * 12);
    float* vj = (float*)malloc((n ? n : 1) * 4);
    float* wi = (float*)malloc((I ? I : 1) * 4);
    for (size_t i = 0; i < I; ++i)
        memcpy(x + 3 * i, xyz + 3 * (size_t)initial[i], 12);
    memcpy(xp, x, I * 12);
    for (size_t j = 0; j < n; ++j)
        vj[j] = 1.0f;
    for (size_t i = 0; i < I; ++i)
        wi[i] = 1.0f;
    oracle_cloud* cp = oracle_cloud_create(xyz, n, NULL, 0, 0);
    if (uniform) /* :371-381 */
        for (size_t j = 0; j < n; ++j)
            vj[j] = wlop_density(cp, xyz + 3 * j, h, h4sq);
    for (size_t it = 0; it < K; ++it)
    {
        oracle_cloud* cq = oracle_cloud_create(x, I, NULL, 0, 0); /* :385-389 */
        if (uniform) /* :391-401 */
            for (size_t i = 0; i < I; ++i)
                wi[i] = wlop_density(cq, x + 3 * i, h, h4sq);
        for (size_t i = 0; i < I; ++i)
        {
            const float* q = x + 3 * i;
            /* solve_first_energy_median, :106-168 */
            u32vec_t nb = ball(cp, q, h);
            float sum = 0.f, med[3] = {0.f, 0.f, 0.f};
            for (size_t j = 0; j < nb.n; ++j)
            {
                const float* p = xyz + 3 * (size_t)nb.a[j];
                if (same_point(q, p))
                    continue;
                float r2    = sqdist3(q, p);
                float r     = sqrtf(r2);
                float v     = vj[nb.a[j]];
                float alpha = fp_equals(r, zero, eps) ? zero : theta_f(r2, h4sq) / r;
                float coeff = fp_equals(v, zero, eps) ? zero : alpha / v;
                for (int a = 0; a < 3; ++a)
                    med[a] = med[a] + coeff * p[a];
                sum += coeff;
            }
            free(nb.a);
            if (fp_equals(sum, zero, eps))
                med[0] = q[0], med[1] = q[1], med[2] = q[2];
            else
                med[0] = med[0] / sum, med[1] = med[1] / sum, med[2] = med[2] / sum;
            /* solve_second_energy_repulsion_force, :170-224 */
            nb = ball(cq, q, h);
            float rsum = 0.f, rep[3] = {0.f, 0.f, 0.f};
            for (size_t j = 0; j < nb.n; ++j)
            {
                const float* qi = x + 3 * (size_t)nb.a[j];
                if (same_point(qi, q))
                    continue;
                float d[3]  = {q[0] - qi[0], q[1] - qi[1], q[2] - qi[2]};
                float r2    = sqdist3(q, qi);
                float r     = sqrtf(r2);
                float beta  = fp_equals(r, zero, eps) ? zero : theta_f(r2, h4sq) / r;
                float coeff = wi[nb.a[j]] * beta;
                for (int a = 0; a < 3; ++a)
                    rep[a] = rep[a] + coeff * d[a];
                rsum += coeff;
            }
            free(nb.a);
            float s = fp_equals(rsum, zero, eps) ? zero : mu / rsum;
            for (int a = 0; a < 3; ++a)
                xp[3 * i + a] = med[a] + s * rep[a]; /* :428-431 */
        }
        oracle_cloud_destroy(cq);
        memcpy(x, xp, I * 12); /* :434 */
    }
    memcpy(out, xp, I * 12);
    oracle_cloud_destroy(cp);
    free(x), free(xp), free(vj), free(wi);
}

/* ------------------------------------------------------------------------------------------
 * propagate_normal_orientations: algorithm/estimate_normals.hpp:187-302
 *
 * knn: n rows of k neighbour indices (nearest first, -1 = none) — the directed kNN graph of
 * graph/knn_adjacency_list.hpp:117-156.  Root = the FIRST point of maximal z
 * (std::max_element, :223-230); its normal is replaced by (0, 0, 1) (:233-237).  Breadth-first
 * search as graph/search.hpp:41-85 (FIFO queue; a vertex is marked when first reached and the
 * edge that reaches it is the only one whose op runs): the reached normal is negated iff
 * inner_product(n_parent, n_child) < 0 and |inner_product| >= 1e-5 (:289-300,
 * common/norm.hpp:34-45, common/vector3d_queries.hpp:31-35), the parent's normal being its
 * final one.  Vertices the search never reaches keep their normals.
 *
 * Edge order: the reference keeps edges in a std::unordered_multimap keyed by the source vertex
 * (graph/directed_adjacency_list.hpp:79-83) and walks equal_range(source) (:183); the order of
 * equal keys is the standard library's.  libstdc++ (this image, and the only build of the
 * reference that can be run here) yields them in REVERSE insertion order, i.e. furthest
 * neighbour first: reverse_edges = 1.  reverse_edges = 0 is insertion order (nearest first).
 *
 * Parity status: PINNED against the unmodified reference compiled here
 * (oracle/ref_bridge_orient.cpp) with reverse_edges = 1.
 * ---------------------------------------------------------------------------------------- */
void oracle_propagate_normal_orientations(const float* xyz, size_t n, const int64_t* knn, size_t k,
                                          int reverse_edges, float* normals)
{
    if (n == 0)
        return;
    size_t root = 0;
    for (size_t i = 1; i < n; ++i)
        if (xyz[3 * root + 2] < xyz[3 * i + 2])
            root = i;
    normals[3 * root] = 0.f, normals[3 * root + 1] = 0.f, normals[3 * root + 2] = 1.f;
    uint8_t* visited = (uint8_t*)calloc(n, 1);
    uint32_t* queue  = (uint32_t*)malloc(n * sizeof(uint32_t) + 4);
    size_t head = 0, tail = 0;
    queue[tail++] = (uint32_t)root;
    while (head < tail)
    {
        size_t const u = queue[head++];
        for (size_t e = 0; e < k; ++e)
        {
            int64_t const v = knn[u * k + (reverse_edges ? k - 1 - e : e)];
            if (v < 0 || visited[v])
                continue;
            const float* n1 = normals + 3 * u;
            float* n2       = normals + 3 * (size_t)v;
            volatile float xx = n2[0] * n1[0], yy = n2[1] * n1[1], zz = n2[2] * n1[2];
            volatile float s  = xx + yy;
            float const prod  = s + zz;
            if (prod < 0.f && !fp_equals(prod, 0.f, (float)1e-5))
                n2[0] = -n2[0], n2[1] = -n2[1], n2[2] = -n2[2];
            visited[v] = 1;
            queue[tail++] = (uint32_t)v;
        }
        visited[u] = 1; /* search.hpp:81-82; only matters for the root, which has no self edge */
    }
    free(visited), free(queue);
}
