/* oracle/geodesic.c -- CPU restatement (TEST INFRASTRUCTURE ONLY, never linked by the product)
 *
 * Geodesic-neighbourhood selection of the reference's tangent_frames():
 *   /root/reference/RVGP/lib/ptu_dijkstra.pyx:361-394  (per-source truncated Dijkstra driver)
 *   /root/reference/RVGP/lib/ptu_dijkstra.pyx:444-690  (Fibonacci node / heap operations)
 *
 * Which K+1 nodes end up in a neighbourhood is decided by the pop order of that heap among
 * equal keys (all edge weights are 1.0), so the heap is restated operation by operation with
 * index links instead of pointers (-1 == NULL).  Keys are the reference's doubles (val + w);
 * they only ever hold small integers, so comparisons are exact either way.
 *
 * Pinned by tests/test_oracle_pinned.py against sequences exported by an instrumented build of the
 * reference (oracle/build_ref.py --instrumented) stored in tests/golden/.
 *
 * Differences from the reference that do not change results:
 *   - nodes touched by a source are reset afterwards from a touched-list instead of
 *     re-initialising all N nodes per source (pyx:364-365 is O(N^2) overall);
 *   - link() is written as a loop (the reference's recursion is a tail call).
 */
#include <stdlib.h>
#include <string.h>

enum { SCANNED = 0, NOT_IN_HEAP = 1, IN_HEAP = 2 }; /* pyx:448-451 */

typedef struct {
    double *val;
    int *rank, *state, *parent, *left, *right, *child;
    int min_node;
    int roots_by_rank[100]; /* pyx:573 */
} heap_t;

static int rightmost(heap_t *h, int n) { while (h->right[n] >= 0) n = h->right[n]; return n; } /* pyx:485-490 */
static int leftmost(heap_t *h, int n) { while (h->left[n] >= 0) n = h->left[n]; return n; }   /* pyx:496-501 */

/* pyx:526-536 */
static void add_sibling(heap_t *h, int node, int sib) {
    int t = rightmost(h, node);
    h->right[t] = sib;
    h->left[sib] = t;
    h->right[sib] = -1;
    h->parent[sib] = h->parent[node];
    if (h->parent[sib] >= 0) h->rank[h->parent[sib]] += 1;
}

/* pyx:507-520 */
static void add_child(heap_t *h, int node, int ch) {
    h->parent[ch] = node;
    if (h->child[node] >= 0) {
        add_sibling(h, h->child[node], ch);
    } else {
        h->child[node] = ch;
        h->right[ch] = -1;
        h->left[ch] = -1;
        h->rank[node] = 1;
    }
}

/* pyx:542-560 : note parent.children is re-pointed whether or not `node` was the pointee */
static void remove_node(heap_t *h, int node) {
    int p = h->parent[node];
    if (p >= 0) {
        h->rank[p] -= 1;
        if (h->left[node] >= 0) h->child[p] = h->left[node];
        else if (h->right[node] >= 0) h->child[p] = h->right[node];
        else h->child[p] = -1;
    }
    if (h->left[node] >= 0) h->right[h->left[node]] = h->right[node];
    if (h->right[node] >= 0) h->left[h->right[node]] = h->left[node];
    h->left[node] = -1;
    h->right[node] = -1;
    h->parent[node] = -1;
}

/* pyx:579-589 */
static void insert_node(heap_t *h, int node) {
    if (h->min_node >= 0) {
        add_sibling(h, h->min_node, node);
        if (h->val[node] < h->val[h->min_node]) h->min_node = node;
    } else {
        h->min_node = node;
    }
}

/* pyx:595-608 */
static void decrease_val(heap_t *h, int node, double newval) {
    h->val[node] = newval;
    if (h->parent[node] >= 0 && h->val[h->parent[node]] >= newval) {
        remove_node(h, node);
        insert_node(h, node);
    } else if (h->val[h->min_node] > h->val[node]) {
        h->min_node = node;
    }
}

/* pyx:614-636 */
static void link_node(heap_t *h, int node) {
    for (;;) {
        int r = h->rank[node];
        if (h->roots_by_rank[r] < 0) { h->roots_by_rank[r] = node; return; }
        int ln = h->roots_by_rank[r];
        h->roots_by_rank[r] = -1;
        if (h->val[node] < h->val[ln] || node == h->min_node) {
            remove_node(h, ln);
            add_child(h, node, ln);
        } else {
            remove_node(h, node);
            add_child(h, ln, node);
            node = ln;
        }
    }
}

/* pyx:642-690 */
static int remove_min(heap_t *h) {
    int temp, temp_right, out, i;
    int mn = h->min_node;
    if (h->child[mn] >= 0) {
        temp = leftmost(h, h->child[mn]);
        while (temp >= 0) {
            temp_right = h->right[temp];
            remove_node(h, temp);
            add_sibling(h, mn, temp);
            temp = temp_right;
        }
        h->child[mn] = -1;
    }
    temp = leftmost(h, mn);
    if (temp == mn) {
        if (h->right[mn] >= 0) temp = h->right[mn];
        else { out = mn; h->min_node = -1; return out; }
    }
    out = mn;
    remove_node(h, mn);
    h->min_node = temp;
    for (i = 0; i < 100; ++i) h->roots_by_rank[i] = -1;
    while (temp >= 0) {
        if (h->val[temp] < h->val[h->min_node]) h->min_node = temp;
        temp_right = h->right[temp];
        link_node(h, temp);
        temp = temp_right;
    }
    return out;
}

/* pyx:361-394.  seq is (N, K+1) int32 row-major; counts[i] = number of pops of source i.
 * Rows whose component is smaller than K+1 keep the previous source's tail (pyx:350: the index
 * buffer is allocated once outside the loop); for source 0 the tail is zero here (np.empty there).
 * weights may be NULL (all 1.0).  Returns 0, or -1 on allocation failure (pyx:358-359). */
int oracle_geodesic_neighbourhoods(const int *indptr, const int *indices, const double *weights,
                                   int N, int K, int *seq, int *counts) {
    heap_t h;
    int Kp1 = K + 1, i, k, ntouched;
    int *touched = (int *)malloc(sizeof(int) * (size_t)N);
    int *cur = (int *)calloc((size_t)Kp1, sizeof(int));
    h.val = (double *)malloc(sizeof(double) * (size_t)N);
    h.rank = (int *)malloc(sizeof(int) * (size_t)N);
    h.state = (int *)malloc(sizeof(int) * (size_t)N);
    h.parent = (int *)malloc(sizeof(int) * (size_t)N);
    h.left = (int *)malloc(sizeof(int) * (size_t)N);
    h.right = (int *)malloc(sizeof(int) * (size_t)N);
    h.child = (int *)malloc(sizeof(int) * (size_t)N);
    if (!touched || !cur || !h.val || !h.rank || !h.state || !h.parent || !h.left || !h.right || !h.child) return -1;
    for (k = 0; k < N; ++k) { /* initialize_node, pyx:465-479 */
        h.val[k] = 0; h.rank[k] = 0; h.state[k] = NOT_IN_HEAP;
        h.parent[k] = h.left[k] = h.right[k] = h.child[k] = -1;
    }
    for (i = 0; i < N; ++i) {
        int scanned = 0;
        ntouched = 0;
        h.min_node = -1;
        insert_node(&h, i);
        touched[ntouched++] = i;
        while (h.min_node >= 0 && scanned <= K) {
            int v = remove_min(&h);
            h.state[v] = SCANNED;
            cur[scanned] = v;
            scanned += 1;
            if (scanned <= K) {
                for (k = indptr[v]; k < indptr[v + 1]; ++k) {
                    int c = indices[k];
                    if (h.state[c] != SCANNED) {
                        double next_val = h.val[v] + (weights ? weights[k] : 1.0);
                        if (h.state[c] == NOT_IN_HEAP) {
                            h.state[c] = IN_HEAP;
                            h.val[c] = next_val;
                            insert_node(&h, c);
                            touched[ntouched++] = c;
                        } else if (h.val[c] > next_val) {
                            decrease_val(&h, c, next_val);
                        }
                    }
                }
            }
        }
        counts[i] = scanned;
        memcpy(seq + (size_t)i * Kp1, cur, sizeof(int) * (size_t)Kp1);
        for (k = 0; k < ntouched; ++k) {
            int t = touched[k];
            h.val[t] = 0; h.rank[t] = 0; h.state[t] = NOT_IN_HEAP;
            h.parent[t] = h.left[t] = h.right[t] = h.child[t] = -1;
        }
    }
    free(touched); free(cur); free(h.val); free(h.rank); free(h.state);
    free(h.parent); free(h.left); free(h.right); free(h.child);
    return 0;
}
