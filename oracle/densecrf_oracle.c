/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * Single-threaded C restatement of the dense-CRF mean-field inference that the reference reaches through
 * pydensecrf (utils.py:74-91: DenseCRF2D / setUnaryEnergy / addPairwiseGaussian / addPairwiseBilateral /
 * inference).  pydensecrf (Cython over Kraehenbuehl's densecrf, unpinned in README.md:43) is NOT vendored under
 * /root/reference and not installable here, so this restates the published algorithm:
 *   P. Kraehenbuehl, V. Koltun, "Efficient Inference in Fully Connected CRFs with Gaussian Edge Potentials", 2011
 *   A. Adams, J. Baek, A. Davis, "Fast High-Dimensional Filtering Using the Permutohedral Lattice", 2010
 * following densecrf's permutohedral.cpp (elevate with scale_i = (d+1)*sqrt(2/3)/sqrt((i+1)(i+2)), round to the
 * nearest 0-coloured simplex, rank, barycentric weights, splat / blur (axes 0..d, 0.5 neighbour weight) / slice
 * with alpha = 1/(1+2^-d)), DenseKernel with DIAG_KERNEL + NORMALIZE_SYMMETRIC (norm = 1/sqrt(K1 + 1e-20)),
 * PottsCompatibility (message -w*K(Q)), and DenseCRF::inference (Q = softmax(-U); Q = softmax(-U + sum w K(Q))).
 *
 * Parity status: UNPINNED by the reference (no golden CRF outputs exist).  Pinned by analytic checks in
 * tests/test_oracle.py: barycentric weights >= 0 summing to 1, constant-image K1 gain, brute-force Gaussian bound.
 *
 * Layout: unary / Q are [M, N] label-major float32 (the numpy layout pydensecrf takes), N = H*W, pixel = y*W + x.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 5

typedef struct {
  int d, N, M;             /* feature dim, points, lattice points */
  int* offset;             /* [N*(d+1)] lattice index per simplex vertex */
  float* bary;             /* [N*(d+1)] */
  int* n1;                 /* [(d+1)*M] blur neighbours, -1 = missing */
  int* n2;
} lattice_t;

/* ---------- hash table on int16 keys of length d (densecrf HashTable) ---------- */
typedef struct {
  int d, cap, filled;
  short* keys;   /* [cap_keys * d] in insertion order */
  int* table;    /* [cap] -> key index or -1 */
} hash_t;

static size_t hash_key(const short* k, int d) {
  size_t r = 0;
  for (int i = 0; i < d; i++) { r += (size_t)(long)k[i]; r *= 1664525; }
  return r;
}
static void hash_init(hash_t* h, int d, int n_elements) {
  h->d = d; h->cap = 2 * n_elements; h->filled = 0;
  h->keys = (short*)malloc(sizeof(short) * (size_t)(h->cap / 2 + 10) * d);
  h->table = (int*)malloc(sizeof(int) * (size_t)h->cap);
  for (int i = 0; i < h->cap; i++) h->table[i] = -1;
}
static int hash_find(hash_t* h, const short* k, int create) {
  size_t c = hash_key(k, h->d) % (size_t)h->cap;
  for (;;) {
    int e = h->table[c];
    if (e == -1) {
      if (!create) return -1;
      memcpy(h->keys + (size_t)h->filled * h->d, k, sizeof(short) * h->d);
      h->table[c] = h->filled;
      return h->filled++;
    }
    if (memcmp(h->keys + (size_t)e * h->d, k, sizeof(short) * h->d) == 0) return e;
    c++;
    if (c == (size_t)h->cap) c = 0;
  }
}
static void hash_free(hash_t* h) { free(h->keys); free(h->table); }

/* ---------- Permutohedral::init ---------- */
void oracle_lattice_free(lattice_t* L) {
  if (!L) return;
  free(L->offset); free(L->bary); free(L->n1); free(L->n2); free(L);
}

lattice_t* oracle_lattice_build(const float* feature /* [N, d] */, int d, int N) {
  lattice_t* L = (lattice_t*)calloc(1, sizeof(lattice_t));
  L->d = d; L->N = N;
  L->offset = (int*)malloc(sizeof(int) * (size_t)N * (d + 1));
  L->bary = (float*)malloc(sizeof(float) * (size_t)N * (d + 1));
  hash_t H;
  hash_init(&H, d, N * (d + 1));
  float scale_factor[MAXD];
  float elevated[MAXD + 1], barycentric[MAXD + 2];
  int rem0[MAXD + 1], rank[MAXD + 1];
  short canonical[(MAXD + 1) * (MAXD + 1)], key[MAXD + 1];
  for (int i = 0; i <= d; i++) {
    for (int j = 0; j <= d - i; j++) canonical[i * (d + 1) + j] = (short)i;
    for (int j = d - i + 1; j <= d; j++) canonical[i * (d + 1) + j] = (short)(i - (d + 1));
  }
  float inv_std_dev = sqrtf(2.0f / 3.0f) * (d + 1);
  for (int i = 0; i < d; i++) scale_factor[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);
  for (int k = 0; k < N; k++) {
    const float* f = feature + (size_t)k * d;
    float sm = 0;
    for (int j = d; j > 0; j--) {
      float cf = f[j - 1] * scale_factor[j - 1];
      elevated[j] = sm - j * cf;
      sm += cf;
    }
    elevated[0] = sm;
    float down_factor = 1.0f / (d + 1);
    float up_factor = (float)(d + 1);
    int sum = 0;
    for (int i = 0; i <= d; i++) {
      int rd2;
      float v = down_factor * elevated[i];
      float up = ceilf(v) * up_factor;
      float down = floorf(v) * up_factor;
      if (up - elevated[i] < elevated[i] - down) rd2 = (short)up;
      else rd2 = (short)down;
      rem0[i] = rd2;
      sum += (int)(rd2 * down_factor);
    }
    for (int i = 0; i <= d; i++) rank[i] = 0;
    for (int i = 0; i < d; i++) {
      float di = elevated[i] - rem0[i];
      for (int j = i + 1; j <= d; j++)
        if (di < elevated[j] - rem0[j]) rank[i]++;
        else rank[j]++;
    }
    for (int i = 0; i <= d; i++) {
      rank[i] += sum;
      if (rank[i] < 0) { rank[i] += d + 1; rem0[i] += d + 1; }
      else if (rank[i] > d) { rank[i] -= d + 1; rem0[i] -= d + 1; }
    }
    for (int i = 0; i <= d + 1; i++) barycentric[i] = 0;
    for (int i = 0; i <= d; i++) {
      float v = (elevated[i] - rem0[i]) * down_factor;
      barycentric[d - rank[i]] += v;
      barycentric[d - rank[i] + 1] -= v;
    }
    barycentric[0] += 1.0f + barycentric[d + 1];
    for (int remainder = 0; remainder <= d; remainder++) {
      for (int i = 0; i < d; i++) key[i] = (short)(rem0[i] + canonical[remainder * (d + 1) + rank[i]]);
      L->offset[(size_t)k * (d + 1) + remainder] = hash_find(&H, key, 1);
      L->bary[(size_t)k * (d + 1) + remainder] = barycentric[remainder];
    }
  }
  int M = H.filled;
  L->M = M;
  L->n1 = (int*)malloc(sizeof(int) * (size_t)(d + 1) * M);
  L->n2 = (int*)malloc(sizeof(int) * (size_t)(d + 1) * M);
  short n1[MAXD + 1], n2[MAXD + 1];
  for (int j = 0; j <= d; j++) {
    for (int i = 0; i < M; i++) {
      const short* kk = H.keys + (size_t)i * d;
      for (int k = 0; k < d; k++) { n1[k] = (short)(kk[k] - 1); n2[k] = (short)(kk[k] + 1); }
      if (j < d) { n1[j] = (short)(kk[j] + d); n2[j] = (short)(kk[j] - d); }
      L->n1[(size_t)j * M + i] = hash_find(&H, n1, 0);
      L->n2[(size_t)j * M + i] = hash_find(&H, n2, 0);
    }
  }
  hash_free(&H);
  return L;
}

int oracle_lattice_size(const lattice_t* L) { return L->M; }
void oracle_lattice_get(const lattice_t* L, int* offset, float* bary) {
  memcpy(offset, L->offset, sizeof(int) * (size_t)L->N * (L->d + 1));
  memcpy(bary, L->bary, sizeof(float) * (size_t)L->N * (L->d + 1));
}

/* ---------- Permutohedral::seqCompute: out[N, vs] = slice(blur(splat(in[N, vs]))) ---------- */
void oracle_lattice_compute(const lattice_t* L, float* out, const float* in, int vs) {
  int d = L->d, N = L->N, M = L->M;
  float* values = (float*)calloc((size_t)(M + 2) * vs, sizeof(float));
  float* new_values = (float*)calloc((size_t)(M + 2) * vs, sizeof(float));
  for (int i = 0; i < N; i++)
    for (int j = 0; j <= d; j++) {
      int o = L->offset[(size_t)i * (d + 1) + j] + 1;
      float w = L->bary[(size_t)i * (d + 1) + j];
      for (int k = 0; k < vs; k++) values[(size_t)o * vs + k] += w * in[(size_t)i * vs + k];
    }
  for (int j = 0; j <= d; j++) {
    for (int i = 0; i < M; i++) {
      float* old_val = values + (size_t)(i + 1) * vs;
      float* new_val = new_values + (size_t)(i + 1) * vs;
      int a = L->n1[(size_t)j * M + i] + 1, b = L->n2[(size_t)j * M + i] + 1;
      float* n1v = values + (size_t)a * vs;
      float* n2v = values + (size_t)b * vs;
      for (int k = 0; k < vs; k++) new_val[k] = old_val[k] + 0.5f * (n1v[k] + n2v[k]);
    }
    float* t = values; values = new_values; new_values = t;
  }
  float alpha = 1.0f / (1 + powf(2, -d));
  for (int i = 0; i < N; i++) {
    for (int k = 0; k < vs; k++) out[(size_t)i * vs + k] = 0;
    for (int j = 0; j <= d; j++) {
      int o = L->offset[(size_t)i * (d + 1) + j] + 1;
      float w = L->bary[(size_t)i * (d + 1) + j];
      for (int k = 0; k < vs; k++) out[(size_t)i * vs + k] += w * values[(size_t)o * vs + k] * alpha;
    }
  }
  free(values); free(new_values);
}

/* features exactly as DenseCRF2D::addPairwiseGaussian / addPairwiseBilateral build them */
void oracle_features_gaussian(float* f /* [N,2] */, int H, int W, float sx, float sy) {
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      f[((size_t)j * W + i) * 2 + 0] = i / sx;
      f[((size_t)j * W + i) * 2 + 1] = j / sy;
    }
}
void oracle_features_bilateral(float* f /* [N,5] */, const uint8_t* im, int H, int W, float sx, float sy, float sr,
                               float sg, float sb) {
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      size_t p = (size_t)j * W + i;
      f[p * 5 + 0] = i / sx;
      f[p * 5 + 1] = j / sy;
      f[p * 5 + 2] = im[p * 3 + 0] / sr;
      f[p * 5 + 3] = im[p * 3 + 1] / sg;
      f[p * 5 + 4] = im[p * 3 + 2] / sb;
    }
}

static void exp_and_normalize(float* Q, const float* in, int N, int M) { /* pixel-major [N, M] */
  for (int i = 0; i < N; i++) {
    const float* b = in + (size_t)i * M;
    float* q = Q + (size_t)i * M;
    float mx = b[0];
    for (int k = 1; k < M; k++) if (b[k] > mx) mx = b[k];
    float s = 0;
    for (int k = 0; k < M; k++) { q[k] = expf(b[k] - mx); s += q[k]; }
    for (int k = 0; k < M; k++) q[k] /= s;
  }
}

/* DenseCRF::inference with the reference's two pairwise terms (either may be disabled with compat == 0).
 * timings_ms (optional, [2]): lattice build, mean-field loop. */
int oracle_crf_inference(int H, int W, int M, int iters, const float* unary /* [M,N] */, const uint8_t* image,
                         float sxy_g, float compat_g, float sxy_b, float srgb_b, float compat_b,
                         float* Q_out /* [M,N] */) {
  int N = H * W;
  lattice_t* Lg = NULL; lattice_t* Lb = NULL;
  float* norm_g = NULL; float* norm_b = NULL;
  float* ones = (float*)malloc(sizeof(float) * N);
  for (int i = 0; i < N; i++) ones[i] = 1.0f;
  if (compat_g != 0) {
    float* f = (float*)malloc(sizeof(float) * (size_t)N * 2);
    oracle_features_gaussian(f, H, W, sxy_g, sxy_g);
    Lg = oracle_lattice_build(f, 2, N);
    free(f);
    norm_g = (float*)malloc(sizeof(float) * N);
    oracle_lattice_compute(Lg, norm_g, ones, 1);
    for (int i = 0; i < N; i++) norm_g[i] = 1.0f / sqrtf(norm_g[i] + 1e-20f);
  }
  if (compat_b != 0) {
    float* f = (float*)malloc(sizeof(float) * (size_t)N * 5);
    oracle_features_bilateral(f, image, H, W, sxy_b, sxy_b, srgb_b, srgb_b, srgb_b);
    Lb = oracle_lattice_build(f, 5, N);
    free(f);
    norm_b = (float*)malloc(sizeof(float) * N);
    oracle_lattice_compute(Lb, norm_b, ones, 1);
    for (int i = 0; i < N; i++) norm_b[i] = 1.0f / sqrtf(norm_b[i] + 1e-20f);
  }
  free(ones);
  size_t NM = (size_t)N * M;
  float* U = (float*)malloc(sizeof(float) * NM);      /* pixel-major -unary */
  float* Q = (float*)malloc(sizeof(float) * NM);
  float* tmp1 = (float*)malloc(sizeof(float) * NM);
  float* tin = (float*)malloc(sizeof(float) * NM);
  float* tout = (float*)malloc(sizeof(float) * NM);
  for (int i = 0; i < N; i++)
    for (int k = 0; k < M; k++) U[(size_t)i * M + k] = -unary[(size_t)k * N + i];
  exp_and_normalize(Q, U, N, M);
  for (int it = 0; it < iters; it++) {
    memcpy(tmp1, U, sizeof(float) * NM);
    for (int term = 0; term < 2; term++) {
      lattice_t* L = term == 0 ? Lg : Lb;
      const float* norm = term == 0 ? norm_g : norm_b;
      float w = term == 0 ? compat_g : compat_b;
      if (!L) continue;
      for (int i = 0; i < N; i++)
        for (int k = 0; k < M; k++) tin[(size_t)i * M + k] = Q[(size_t)i * M + k] * norm[i];
      oracle_lattice_compute(L, tout, tin, M);
      /* Potts: message = -w * (norm * K(norm * Q)); tmp1 -= message */
      for (int i = 0; i < N; i++)
        for (int k = 0; k < M; k++) tmp1[(size_t)i * M + k] -= -w * (tout[(size_t)i * M + k] * norm[i]);
    }
    exp_and_normalize(Q, tmp1, N, M);
  }
  for (int i = 0; i < N; i++)
    for (int k = 0; k < M; k++) Q_out[(size_t)k * N + i] = Q[(size_t)i * M + k];
  free(U); free(Q); free(tmp1); free(tin); free(tout);
  free(norm_g); free(norm_b);
  oracle_lattice_free(Lg); oracle_lattice_free(Lb);
  return 0;
}
