// host_ecc.cu -- host layer above the decoder ABI (include/ldpc_b200_ecc.h): matrix file readers, the ECC record
// of mkLDPC, the encoder, and one tester point.  C++ on the host; the only device code here is the encoder and
// the random-message generator used by ldpc_b200_ecc_run.
#include "common.cuh"
#include "philox.cuh"
#include "../../include/ldpc_b200_ecc.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using namespace ldpc;

// ------------------------------------------------------------------------------------------------------------------
// matrices
// ------------------------------------------------------------------------------------------------------------------

typedef std::vector<uint32_t> Big;   // little-endian 32-bit limbs of a circulant word (arbitrary precision, .q files)

struct ldpc_b200_matrix {
    int rows = 0, cols = 0;
    std::vector<int32_t> row_ptr, col_idx;
    int sz = 0, brows = 0, bcols = 0;          // quasi-cyclic description when sz > 0
    std::vector<Big> words;                    // brows*bcols circulant words
};

static Big parse_decimal(const std::string &t)
{
    Big v(1, 0);
    for (char ch : t) {
        if (ch < '0' || ch > '9') continue;
        uint64_t carry = (uint64_t)(ch - '0');
        for (size_t i = 0; i < v.size(); i++) {
            uint64_t x = (uint64_t)v[i] * 10u + carry;
            v[i] = (uint32_t)x;
            carry = x >> 32;
        }
        if (carry) v.push_back((uint32_t)carry);
    }
    return v;
}
static bool big_bit(const Big &v, int k) { return (size_t)(k >> 5) < v.size() && ((v[k >> 5] >> (k & 31)) & 1u); }
static int big_bits(const Big &v)
{
    for (int i = (int)v.size() - 1; i >= 0; i--)
        if (v[i]) return i * 32 + (32 - __builtin_clz(v[i]));
    return 0;
}
static int big_popcount(const Big &v)
{
    int n = 0;
    for (uint32_t w : v) n += __builtin_popcount(w);
    return n;
}

// QuasiCyclic.hs:19-25: entry (m*sz+i, n*sz+j) is set iff bit ((j-i) mod sz) of word (m,n) is set.
static void expand_qc(ldpc_b200_matrix *m)
{
    int sz = m->sz;
    m->rows = m->brows * sz;
    m->cols = m->bcols * sz;
    m->row_ptr.assign(1, 0);
    m->col_idx.clear();
    std::vector<int> ks;
    for (int br = 0; br < m->brows; br++)
        for (int i = 0; i < sz; i++) {
            for (int bc = 0; bc < m->bcols; bc++) {
                const Big &w = m->words[(size_t)br * m->bcols + bc];
                ks.clear();
                int nb = std::min(big_bits(w), sz);
                for (int k = 0; k < nb; k++)
                    if (big_bit(w, k)) ks.push_back(bc * sz + (i + k) % sz);
                std::sort(ks.begin(), ks.end());
                m->col_idx.insert(m->col_idx.end(), ks.begin(), ks.end());
            }
            m->row_ptr.push_back((int32_t)m->col_idx.size());
        }
}

static int parse_q(const std::string &text, ldpc_b200_matrix *m)
{
    // QuasiCyclic.hs:52-56: first token = circulant size, then a Matlab-style integer matrix (one block row per line)
    std::istringstream in(text);
    std::string line;
    std::string first;
    if (!(in >> first)) { set_error("q: empty file"); return LDPC_B200_EINVAL; }
    m->sz = atoi(first.c_str());
    if (m->sz <= 0) { set_error("q: bad circulant size '%s'", first.c_str()); return LDPC_B200_EINVAL; }
    std::getline(in, line);   // rest of the first line
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok;
        int n = 0;
        while (ls >> tok) { m->words.push_back(parse_decimal(tok)); n++; }
        if (n == 0) continue;
        if (m->bcols == 0) m->bcols = n;
        else if (n != m->bcols) { set_error("q: ragged block row %d (%d entries, expected %d)", m->brows, n, m->bcols); return LDPC_B200_EINVAL; }
        m->brows++;
    }
    if (m->brows == 0) { set_error("q: no block rows"); return LDPC_B200_EINVAL; }
    expand_qc(m);
    return 0;
}

static int parse_alist(const std::string &text, ldpc_b200_matrix *m)
{
    // Alist.hs:30-46: every zero is dropped from the token stream first; "nrows ncols", two ignored numbers, nrows row
    // weights, ncols column weights, the row lists (1-based columns); the column lists are read and discarded.
    std::vector<long> t;
    {
        std::istringstream in(text);
        long v;
        while (in >> v) if (v != 0) t.push_back(v);
    }
    size_t p = 0;
    auto item = [&](long &v) -> bool { if (p >= t.size()) return false; v = t[p++]; return true; };
    long n, mm, skip;
    if (!item(n) || !item(mm) || !item(skip) || !item(skip) || n <= 0 || mm <= 0) { set_error("alist: bad header"); return LDPC_B200_EINVAL; }
    std::vector<long> rw(n), cw(mm);
    for (long i = 0; i < n; i++) if (!item(rw[i])) { set_error("alist: truncated row weights"); return LDPC_B200_EINVAL; }
    for (long i = 0; i < mm; i++) if (!item(cw[i])) { set_error("alist: truncated column weights"); return LDPC_B200_EINVAL; }
    m->rows = (int)n; m->cols = (int)mm;
    m->row_ptr.assign(1, 0);
    std::vector<int> ks;
    for (long i = 0; i < n; i++) {
        ks.clear();
        for (long c = 0; c < rw[i]; c++) {
            long v;
            if (!item(v) || v < 1 || v > mm) { set_error("alist: bad entry in row %ld", i); return LDPC_B200_EINVAL; }
            ks.push_back((int)v - 1);
        }
        std::sort(ks.begin(), ks.end());
        ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
        m->col_idx.insert(m->col_idx.end(), ks.begin(), ks.end());
        m->row_ptr.push_back((int32_t)m->col_idx.size());
    }
    return 0;
}

static int parse_m(const std::string &text, ldpc_b200_matrix *m)
{
    // BitMatrix/Matlab.hs:19-26: lines of space separated 0/1
    std::istringstream in(text);
    std::string line;
    m->row_ptr.assign(1, 0);
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok;
        int j = 0;
        bool any = false;
        while (ls >> tok) {
            any = true;
            if (tok == "1") m->col_idx.push_back(j);
            else if (tok != "0") { set_error("m: entry '%s' is not 0/1", tok.c_str()); return LDPC_B200_EINVAL; }
            j++;
        }
        if (!any) continue;
        if (m->cols == 0) m->cols = j;
        else if (j != m->cols) { set_error("m: ragged row %d", m->rows); return LDPC_B200_EINVAL; }
        m->rows++;
        m->row_ptr.push_back((int32_t)m->col_idx.size());
    }
    if (m->rows == 0) { set_error("m: empty matrix"); return LDPC_B200_EINVAL; }
    return 0;
}

extern "C" int ldpc_b200_matrix_parse(const char *kind, const char *text, ldpc_b200_matrix_t **out)
{
    if (!kind || !text || !out) { set_error("matrix_parse: bad arguments"); return LDPC_B200_EINVAL; }
    ldpc_b200_matrix *m = new ldpc_b200_matrix();
    int rc;
    std::string k(kind);
    if (k == "q") rc = parse_q(text, m);
    else if (k == "alist") rc = parse_alist(text, m);
    else if (k == "m") rc = parse_m(text, m);
    else { set_error("matrix_parse: unknown kind '%s'", kind); rc = LDPC_B200_EINVAL; }
    if (rc) { delete m; return rc; }
    *out = m;
    return 0;
}

extern "C" int ldpc_b200_matrix_load(const char *codes_root, const char *path_noext, ldpc_b200_matrix_t **out)
{
    if (!codes_root || !path_noext || !out) { set_error("matrix_load: bad arguments"); return LDPC_B200_EINVAL; }
    // Loader.hs:58-81: suffixes q, alist, m in that order; first existing file wins
    static const char *suffix[3] = {"q", "alist", "m"};
    for (int s = 0; s < 3; s++) {
        std::string p = std::string(codes_root) + "/" + path_noext + "." + suffix[s];
        std::ifstream f(p.c_str());
        if (!f.good()) continue;
        std::stringstream ss;
        ss << f.rdbuf();
        return ldpc_b200_matrix_parse(suffix[s], ss.str().c_str(), out);
    }
    set_error("can not find any matrix files in \"%s\"", path_noext);
    return LDPC_B200_ENOENT;
}

extern "C" void ldpc_b200_matrix_destroy(ldpc_b200_matrix_t *m) { delete m; }

extern "C" int ldpc_b200_matrix_info(const ldpc_b200_matrix_t *m, int *rows, int *cols, int64_t *ones, int *qc_sz)
{
    if (!m) { set_error("matrix_info: null"); return LDPC_B200_EINVAL; }
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (ones) *ones = (int64_t)m->col_idx.size();
    if (qc_sz) *qc_sz = m->sz;
    return 0;
}

extern "C" int ldpc_b200_matrix_csr(const ldpc_b200_matrix_t *m, int32_t *row_ptr, int32_t *col_idx)
{
    if (!m || !row_ptr) { set_error("matrix_csr: bad arguments"); return LDPC_B200_EINVAL; }
    memcpy(row_ptr, m->row_ptr.data(), sizeof(int32_t) * (m->rows + 1));
    if (col_idx && !m->col_idx.empty()) memcpy(col_idx, m->col_idx.data(), sizeof(int32_t) * m->col_idx.size());
    return 0;
}

extern "C" int ldpc_b200_matrix_qc_shifts(const ldpc_b200_matrix_t *m, int *brows, int *bcols, int32_t *shifts)
{
    if (!m) { set_error("matrix_qc_shifts: null"); return LDPC_B200_EINVAL; }
    if (m->sz == 0) { set_error("matrix is not quasi-cyclic"); return LDPC_B200_EUNSUPPORTED; }
    for (const Big &w : m->words) {
        int pc = big_popcount(w);
        // Fast/Arraylet.hs:72-79: decoders on the circulant form need exactly one set bit per non-zero block
        if (pc > 1 || big_bits(w) > m->sz) { set_error("QuasiCyclic matrix has non-powers of two initial value"); return LDPC_B200_EUNSUPPORTED; }
    }
    if (brows) *brows = m->brows;
    if (bcols) *bcols = m->bcols;
    if (shifts)
        for (size_t i = 0; i < m->words.size(); i++) shifts[i] = big_popcount(m->words[i]) ? big_bits(m->words[i]) - 1 : -1;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// ECC record
// ------------------------------------------------------------------------------------------------------------------

struct ldpc_b200_ecc {
    std::string name, decoder_name, code_name, codes_root;
    int algo = 1, precision = 1, max_iters = 0, max_batch = 1;
    int K = 0, N = 0, c_length = 0, num = 0, den = 0;
    ldpc_b200_matrix *H = nullptr, *G = nullptr;
    // device side (lazy)
    ldpc_b200_code_t *dcode = nullptr;
    ldpc_b200_decoder_t *dec = nullptr;
    uint32_t *d_gwords = nullptr;      // QC generator: [g_brows][g_bcols][sz/32] words
    int32_t *d_grow_ptr = nullptr, *d_gcol_idx = nullptr;   // general generator: CSR
    uint32_t *d_msg = nullptr, *d_cw = nullptr, *d_bits = nullptr;
    void *d_llr = nullptr;
    int32_t *d_iters = nullptr;
    uint8_t *d_conv = nullptr;
    uint64_t *d_counters = nullptr;
};

static bool all_digits(const std::string &s)
{
    if (s.empty()) return false;
    for (char c : s) if (c < '0' || c > '9') return false;
    return true;
}

static long gcd_l(long a, long b) { while (b) { long t = a % b; a = b; b = t; } return a; }

extern "C" int ldpc_b200_ecc_open(const char *codes_root, const char *name, int max_batch, ldpc_b200_ecc_t **out)
{
    if (!codes_root || !name || !out || max_batch <= 0) { set_error("ecc_open: bad arguments"); return LDPC_B200_EINVAL; }
    // the matcher of mkLDPC_CodeIO (Utils.hs:100-108): ["ldpc", nm, m, n] or ["ldpc", nm, m, n, x, y], digits only
    std::vector<std::string> xs;
    {
        std::string s(name), cur;
        for (char c : s) { if (c == '/') { xs.push_back(cur); cur.clear(); } else cur += c; }
        xs.push_back(cur);
    }
    static const struct { const char *nm; int algo, prec; } decs[] = {
        {"b200-minsum", LDPC_B200_ALGO_MINSUM, LDPC_B200_PREC_FP32}, {"b200-spa", LDPC_B200_ALGO_SPA, LDPC_B200_PREC_FP32},
        {"b200-minsum-fp64", LDPC_B200_ALGO_MINSUM, LDPC_B200_PREC_FP64}, {"b200-spa-fp64", LDPC_B200_ALGO_SPA, LDPC_B200_PREC_FP64},
        {"b200-spa-cm-fp64", LDPC_B200_ALGO_SPA_CM, LDPC_B200_PREC_FP64},       // the reference's `arraylet-cm` (Fast/CachedMult.hs:229)
        {"b200-arraylet-fp64", LDPC_B200_ALGO_SPA_ARRAYLET, LDPC_B200_PREC_FP64},          // `arraylet` (Fast/Arraylet.hs:135)
        {"b200-arraylet-min-fp64", LDPC_B200_ALGO_MINSUM_ARRAYLET, LDPC_B200_PREC_FP64}};  // `arraylet-min` (Fast/ArrayletMin.hs:136)
    int di = -1;
    if ((xs.size() == 4 || xs.size() == 6) && xs[0] == "ldpc")
        for (int i = 0; i < (int)(sizeof decs / sizeof decs[0]); i++) if (xs[1] == decs[i].nm) di = i;
    bool ok = di >= 0 && all_digits(xs[3]) && (xs.size() == 4 || (all_digits(xs[4]) && all_digits(xs[5])));
    if (!ok) { set_error("'%s' does not match ldpc/<b200-minsum|b200-spa|b200-minsum-fp64|b200-spa-fp64|b200-spa-cm-fp64|b200-arraylet-fp64|b200-arraylet-min-fp64>/<matrix-name>/<max-rounds>[/codeword/message]", name); return LDPC_B200_ENOMATCH; }

    ldpc_b200_ecc *e = new ldpc_b200_ecc();
    e->decoder_name = xs[1]; e->code_name = xs[2]; e->codes_root = codes_root;
    e->algo = decs[di].algo; e->precision = decs[di].prec;
    e->max_iters = atoi(xs[3].c_str()); e->max_batch = max_batch;
    // mkLDPC (Utils.hs:35-51): load G then H, check sizes, rate, transmitted length
    int rc = ldpc_b200_matrix_load(codes_root, (xs[2] + "/G").c_str(), &e->G);
    if (!rc) rc = ldpc_b200_matrix_load(codes_root, (xs[2] + "/H").c_str(), &e->H);
    if (rc) { ldpc_b200_ecc_close(e); return rc; }
    if (e->G->rows + e->G->cols != e->H->cols) {
        set_error("bad code size match(%d,%d)", e->G->rows + e->G->cols, e->H->cols);      // Utils.hs:43-44
        ldpc_b200_ecc_close(e);
        return LDPC_B200_EINVAL;
    }
    e->K = e->G->rows; e->N = e->H->cols;
    long num, den;
    if (xs.size() == 6) { num = atol(xs[4].c_str()); den = atol(xs[5].c_str()); }
    else { num = e->K; den = e->N; }
    if (num <= 0 || den <= 0) { set_error("ecc_open: bad rate %ld/%ld", num, den); ldpc_b200_ecc_close(e); return LDPC_B200_EINVAL; }
    long g = gcd_l(num, den);                     // Data.Ratio keeps fractions reduced (shows up in the name)
    num /= g; den /= g;
    e->num = (int)num; e->den = (int)den;
    e->c_length = (int)(((long)e->K * den) / num);        // Utils.hs:50
    if (e->c_length < e->K || e->c_length > e->N) { set_error("ecc_open: transmitted length %d outside [%d, %d]", e->c_length, e->K, e->N); ldpc_b200_ecc_close(e); return LDPC_B200_EINVAL; }
    char buf[512];
    snprintf(buf, sizeof buf, "ldpc/%s/%s/%d/%d/%d", xs[1].c_str(), xs[2].c_str(), e->max_iters, e->num, e->den);   // Utils.hs:60
    e->name = buf;
    *out = e;
    return 0;
}

extern "C" void ldpc_b200_ecc_close(ldpc_b200_ecc_t *e)
{
    if (!e) return;
    if (e->dec) ldpc_b200_decoder_destroy(e->dec);
    if (e->dcode) ldpc_b200_code_destroy(e->dcode);
    cudaFree(e->d_gwords); cudaFree(e->d_grow_ptr); cudaFree(e->d_gcol_idx); cudaFree(e->d_msg); cudaFree(e->d_cw);
    cudaFree(e->d_bits); cudaFree(e->d_llr); cudaFree(e->d_iters); cudaFree(e->d_conv); cudaFree(e->d_counters);
    delete e->H; delete e->G;
    delete e;
}

extern "C" const char *ldpc_b200_ecc_name(const ldpc_b200_ecc_t *e) { return e ? e->name.c_str() : ""; }
extern "C" int ldpc_b200_ecc_message_length(const ldpc_b200_ecc_t *e) { return e ? e->K : LDPC_B200_EINVAL; }
extern "C" int ldpc_b200_ecc_codeword_length(const ldpc_b200_ecc_t *e) { return e ? e->c_length : LDPC_B200_EINVAL; }
extern "C" int ldpc_b200_ecc_unpunctured_length(const ldpc_b200_ecc_t *e) { return e ? e->N : LDPC_B200_EINVAL; }
extern "C" int ldpc_b200_ecc_max_iters(const ldpc_b200_ecc_t *e) { return e ? e->max_iters : LDPC_B200_EINVAL; }

extern "C" int ldpc_b200_ecc_encode(const ldpc_b200_ecc_t *e, const uint8_t *msg, uint8_t *cw)
{
    if (!e || !msg || !cw) { set_error("ecc_encode: bad arguments"); return LDPC_B200_EINVAL; }
    const ldpc_b200_matrix *G = e->G;
    int P = G->cols;
    std::vector<uint8_t> par(P > 0 ? P : 1, 0);
    if (G->sz > 0) {
        // Fast/Encoder.hs:41-55: res[col] = xor over block rows of mulWord(v_row, G[row,col]); mulWord (:58-63) xors the
        // circulant word rotated left by n for every set bit n of the message word, i.e. message bit n of block row r
        // and generator bit k of block (r,c) toggle parity bit c*sz + (n+k) mod sz.
        int sz = G->sz;
        for (int r = 0; r < G->brows; r++)
            for (int n = 0; n < sz; n++) {
                if (!msg[r * sz + n]) continue;
                for (int c = 0; c < G->bcols; c++) {
                    const Big &w = G->words[(size_t)r * G->bcols + c];
                    int nb = std::min(big_bits(w), sz);
                    for (int k = 0; k < nb; k++)
                        if (big_bit(w, k)) par[c * sz + (n + k) % sz] ^= 1;
                }
            }
    } else {
        // Orig.hs:25-26: v * G over GF(2)
        for (int k = 0; k < G->rows; k++)
            if (msg[k]) for (int q = G->row_ptr[k]; q < G->row_ptr[k + 1]; q++) par[G->col_idx[q]] ^= 1;
    }
    memcpy(cw, msg, e->K);
    for (int i = 0; i < e->c_length - e->K; i++) cw[e->K + i] = par[i];      // Utils.hs:61
    return 0;
}

static int ensure_device(ldpc_b200_ecc *e)
{
    if (e->dec) return 0;
    int rc;
    int br = 0, bc = 0;
    if (e->H->sz > 0 && ldpc_b200_matrix_qc_shifts(e->H, &br, &bc, nullptr) == 0) {
        std::vector<int32_t> sh((size_t)br * bc);
        ldpc_b200_matrix_qc_shifts(e->H, &br, &bc, sh.data());
        rc = ldpc_b200_code_create_qc(e->H->sz, br, bc, sh.data(), &e->dcode);
    } else {
        rc = ldpc_b200_code_create_csr(e->H->rows, e->H->cols, e->H->row_ptr.data(), e->H->col_idx.data(), &e->dcode);
    }
    if (rc) return rc;
    return ldpc_b200_decoder_create(e->dcode, e->algo, e->precision, e->max_batch, &e->dec);
}

extern "C" int ldpc_b200_decode_punctured(ldpc_b200_decoder_t *d, int max_iters, int n_frames, const double *rx, int rx_stride, int take,
                                          uint8_t *msg_bits, int k_take);

extern "C" int ldpc_b200_ecc_decode(ldpc_b200_ecc_t *e, int n_frames, const double *rx, int rx_len, uint8_t *out, uint8_t *flags)
{
    if (!e || n_frames < 0 || (n_frames && (!rx || !out)) || rx_len < 0) { set_error("ecc_decode: bad arguments"); return LDPC_B200_EINVAL; }
    if (n_frames == 0) return 0;
    int rc = ensure_device(e);
    if (rc) return rc;
    int N = e->N, K = e->K;
    int take = std::min(std::min(e->c_length, rx_len), N);
    // unpuncture xs = take c_length xs ++ replicate (N - c_length) 0   (Utils.hs:55); take m_length (Utils.hs:72):
    // the padding zeros are written on the device and only the message bits are copied back
    if (take <= 0) { set_error("ecc_decode: no received values"); return LDPC_B200_EINVAL; }
    rc = ldpc_b200_decode_punctured(e->dec, e->max_iters, n_frames, rx, rx_len, take, out, K);
    if (rc) return rc;
    if (flags) memset(flags, 1, (size_t)n_frames);                     // every decoder returns Just
    return 0;
}

// ---- device encoder + random messages ----

__global__ void k_random_messages(uint32_t *__restrict__ msg, int Wk, int K, int F, uint64_t seed, uint64_t first_frame)
{
    int g = blockIdx.y * blockDim.x + threadIdx.x, f = blockIdx.x;
    if (f >= F || g * 4 >= Wk) return;
    // counter word 1 = 1 separates the message stream from the noise stream (philox.cuh uses 0 there)
    Philox4 r = philox4x32_10((uint32_t)g, 1u, (uint32_t)(first_frame + f), (uint32_t)((first_frame + f) >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int q = 0; q < 4; q++) {
        int w = 4 * g + q;
        if (w >= Wk) break;
        uint32_t v = r.v[q];
        int hi = K - 32 * w;
        if (hi < 32) v &= (hi <= 0) ? 0u : ((1u << hi) - 1u);
        msg[(size_t)f * Wk + w] = v;
    }
}

// Quasi-cyclic generator (Fast/Encoder.hs:41-63). One CTA per frame; thread = one 32-bit word of the parity part.
// gw: [g_brows][g_bcols][WS] words (WS = sz/32). sz must be a multiple of 32.
__global__ void k_encode_qc(const uint32_t *__restrict__ msg, uint32_t *__restrict__ cw, const uint32_t *__restrict__ gw,
                            int sz, int g_brows, int g_bcols, int K, int N, int c_length, int F)
{
    extern __shared__ uint32_t s_msg[];
    int f = blockIdx.x;
    if (f >= F) return;
    int WS = sz / 32, Wk = (K + 31) / 32, Wn = (N + 31) / 32;
    for (int w = threadIdx.x; w < Wk; w += blockDim.x) {
        uint32_t v = msg[(size_t)f * Wk + w];
        s_msg[w] = v;
        cw[(size_t)f * Wn + w] = v;                                    // systematic part
    }
    __syncthreads();
    int pw_total = g_bcols * WS;
    for (int pw = threadIdx.x; pw < pw_total; pw += blockDim.x) {
        int c = pw / WS, w = pw % WS;
        uint32_t acc = 0;
        for (int r = 0; r < g_brows; r++) {
            const uint32_t *g = gw + ((size_t)r * g_bcols + c) * WS;
            for (int vw = 0; vw < WS; vw++) {
                uint32_t v = s_msg[r * WS + vw];
                while (v) {
                    int b = __ffs(v) - 1;
                    v &= v - 1;
                    int n = vw * 32 + b;                               // message bit n of block row r: xor rotl(g, n)
                    // word w of rotl(g, n) = bits [32w - n, 32w - n + 32) of g, cyclically
                    int start = ((32 * w - n) % sz + sz) % sz;
                    int w0 = start >> 5, sh = start & 31;
                    uint32_t lo = g[w0], hi = g[(w0 + 1) % WS];
                    acc ^= __funnelshift_r(lo, hi, sh);
                }
            }
        }
        // parity bit p lives at codeword position K + p; bits at or beyond c_length are not transmitted (Utils.hs:61)
        int p0 = c * sz + 32 * w;
        int keep = c_length - K - p0;
        if (keep < 32) acc &= (keep <= 0) ? 0u : ((1u << keep) - 1u);
        int pos = K + p0;                                               // K and sz are multiples of 32 here
        cw[(size_t)f * Wn + (pos >> 5)] = acc;
    }
    for (int w = Wk + pw_total + threadIdx.x; w < Wn; w += blockDim.x) cw[(size_t)f * Wn + w] = 0;
}

// General generator as CSR rows (Orig.hs:25-26). One CTA per frame, parity accumulated in shared memory.
__global__ void k_encode_csr(const uint32_t *__restrict__ msg, uint32_t *__restrict__ cw, const int32_t *__restrict__ grp,
                             const int32_t *__restrict__ gci, int K, int N, int c_length, int F)
{
    extern __shared__ uint32_t s_cw[];
    int f = blockIdx.x;
    if (f >= F) return;
    int Wk = (K + 31) / 32, Wn = (N + 31) / 32;
    for (int w = threadIdx.x; w < Wn; w += blockDim.x) s_cw[w] = w < Wk ? msg[(size_t)f * Wk + w] : 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        if (!((s_cw[k >> 5] >> (k & 31)) & 1u)) continue;
        for (int q = grp[k]; q < grp[k + 1]; q++) {
            int pos = K + gci[q];
            if (pos < c_length) atomicXor(&s_cw[pos >> 5], 1u << (pos & 31));
        }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < Wn; w += blockDim.x) cw[(size_t)f * Wn + w] = s_cw[w];
}

static int ensure_run_buffers(ldpc_b200_ecc *e)
{
    int rc = ensure_device(e);
    if (rc) return rc;
    if (e->d_counters) return 0;
    size_t B = (size_t)e->max_batch;
    int Wk = (e->K + 31) / 32, Wn = (e->N + 31) / 32;
    const ldpc_b200_matrix *G = e->G;
    if (G->sz > 0 && G->sz % 32 == 0 && e->K % 32 == 0) {
        int WS = G->sz / 32;
        std::vector<uint32_t> gw((size_t)G->brows * G->bcols * WS, 0u);
        for (size_t i = 0; i < G->words.size(); i++)
            for (int w = 0; w < WS && w < (int)G->words[i].size(); w++) gw[i * WS + w] = G->words[i][w];
        LDPC_CUDA(cudaMalloc(&e->d_gwords, gw.size() * sizeof(uint32_t)));
        LDPC_CUDA(cudaMemcpy(e->d_gwords, gw.data(), gw.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    } else {
        LDPC_CUDA(cudaMalloc(&e->d_grow_ptr, sizeof(int32_t) * (G->rows + 1)));
        LDPC_CUDA(cudaMalloc(&e->d_gcol_idx, sizeof(int32_t) * std::max<size_t>(1, G->col_idx.size())));
        LDPC_CUDA(cudaMemcpy(e->d_grow_ptr, G->row_ptr.data(), sizeof(int32_t) * (G->rows + 1), cudaMemcpyHostToDevice));
        if (!G->col_idx.empty()) LDPC_CUDA(cudaMemcpy(e->d_gcol_idx, G->col_idx.data(), sizeof(int32_t) * G->col_idx.size(), cudaMemcpyHostToDevice));
    }
    LDPC_CUDA(cudaMalloc(&e->d_msg, B * Wk * sizeof(uint32_t)));
    LDPC_CUDA(cudaMalloc(&e->d_cw, B * Wn * sizeof(uint32_t)));
    LDPC_CUDA(cudaMalloc(&e->d_bits, B * Wn * sizeof(uint32_t)));
    LDPC_CUDA(cudaMalloc(&e->d_llr, B * e->N * elem_size(e->precision)));
    LDPC_CUDA(cudaMalloc(&e->d_iters, B * sizeof(int32_t)));
    LDPC_CUDA(cudaMalloc(&e->d_conv, B));
    LDPC_CUDA(cudaMalloc(&e->d_counters, 8 * sizeof(uint64_t)));
    return 0;
}

extern "C" int ldpc_b200_ecc_encode_device(ldpc_b200_ecc_t *e, int n_frames, const uint32_t *d_msg_packed, uint32_t *d_cw_packed)
{
    if (!e || n_frames < 0 || (n_frames && (!d_msg_packed || !d_cw_packed))) { set_error("ecc_encode_device: bad arguments"); return LDPC_B200_EINVAL; }
    if (n_frames == 0) return 0;
    int rc = ensure_run_buffers(e);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)ldpc_b200_decoder_stream(e->dec);
    const ldpc_b200_matrix *G = e->G;
    if (e->d_gwords) {
        size_t smem = sizeof(uint32_t) * ((e->K + 31) / 32);
        k_encode_qc<<<n_frames, 128, smem, st>>>(d_msg_packed, d_cw_packed, e->d_gwords, G->sz, G->brows, G->bcols, e->K, e->N, e->c_length, n_frames);
    } else {
        size_t smem = sizeof(uint32_t) * ((e->N + 31) / 32);
        k_encode_csr<<<n_frames, 128, smem, st>>>(d_msg_packed, d_cw_packed, e->d_grow_ptr, e->d_gcol_idx, e->K, e->N, e->c_length, n_frames);
    }
    LDPC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ldpc_b200_ecc_sync(ldpc_b200_ecc_t *e)
{
    if (!e) { set_error("ecc_sync: null"); return LDPC_B200_EINVAL; }
    return e->dec ? ldpc_b200_decoder_sync(e->dec) : 0;
}

extern "C" int ldpc_b200_ecc_run(ldpc_b200_ecc_t *e, double ebn0_db, uint64_t seed, uint64_t first_frame, uint64_t n_frames,
                                 int random_messages, ldpc_b200_counters_t *out)
{
    if (!e || !out) { set_error("ecc_run: bad arguments"); return LDPC_B200_EINVAL; }
    int rc = ensure_device(e);
    if (rc) return rc;
    if (!random_messages)
        return ldpc_b200_simulate(e->dec, e->max_iters, 1, ebn0_db, e->K, e->c_length, seed, first_frame, n_frames, out);
    rc = ensure_run_buffers(e);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)ldpc_b200_decoder_stream(e->dec);
    int Wk = (e->K + 31) / 32;
    LDPC_CUDA(cudaMemsetAsync(e->d_counters, 0, 8 * sizeof(uint64_t), st));
    double ms = 0.0;
    for (uint64_t f0 = 0; f0 < n_frames; f0 += (uint64_t)e->max_batch) {
        int F = (int)std::min<uint64_t>((uint64_t)e->max_batch, n_frames - f0);
        dim3 grid(F, ((Wk + 3) / 4 + 63) / 64);
        k_random_messages<<<grid, 64, 0, st>>>(e->d_msg, Wk, e->K, F, seed, first_frame + f0);
        rc = ldpc_b200_ecc_encode_device(e, F, e->d_msg, e->d_cw);
        if (!rc) rc = ldpc_b200_generate_llr(e->dec, ebn0_db, e->K, e->c_length, seed, first_frame + f0, F, e->d_cw, e->d_llr);
        if (!rc) rc = ldpc_b200_decode_device(e->dec, e->max_iters, 1, F, e->d_llr, e->d_bits, e->d_iters, e->d_conv);
        if (!rc) rc = ldpc_b200_count_errors(e->dec, F, e->K, e->d_bits, e->d_cw, e->d_iters, e->d_conv, e->d_counters);
        if (rc) return rc;
        double k_ms = 0; int64_t nl = 0;
        rc = ldpc_b200_decoder_sync(e->dec);
        if (!rc) rc = ldpc_b200_decoder_last_timing(e->dec, &k_ms, &nl);
        if (rc) return rc;
        ms += k_ms;
    }
    uint64_t h[8];
    LDPC_CUDA(cudaMemcpyAsync(h, e->d_counters, sizeof h, cudaMemcpyDeviceToHost, st));
    LDPC_CUDA(cudaStreamSynchronize(st));
    out->frames += h[0]; out->frame_errors += h[1]; out->bit_errors += h[2]; out->sum_iters += h[3]; out->unconverged += h[4];
    out->kernel_ms += ms;
    return 0;
}

extern "C" int ldpc_b200_ber_line(char *buf, int buf_len, double seconds, const char *name, double ebn0_db, uint64_t frames,
                                  uint64_t bit_errors, int message_length)
{
    // NOTES.txt:3  "39.70 ldpc/model/jpl.1K/200 0.0 256 29891 1.14e-1 +1% -1% 95%": time, name, Eb/N0, frames, bit errors,
    // BER and its 95 % interval relative to the estimate (Wilson score interval on bit_errors / (frames * message_length)).
    double n = (double)frames * (double)message_length;
    double p = n > 0 ? (double)bit_errors / n : 0.0;
    const double z = 1.959963984540054;
    double lo = 0, hi = 0;
    if (n > 0) {
        double den = 1 + z * z / n, ctr = (p + z * z / (2 * n)) / den, half = z * sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den;
        lo = std::max(0.0, ctr - half); hi = std::min(1.0, ctr + half);
    }
    double up = p > 0 ? (hi - p) / p * 100.0 : 0.0, dn = p > 0 ? (p - lo) / p * 100.0 : 0.0;
    // the reference prints the rate like 1.14e-1 (no zero padding of the exponent)
    char ber[32];
    if (p > 0) {
        int ex = (int)floor(log10(p));
        double mant = p / pow(10.0, ex);
        if (mant >= 9.995) { mant /= 10.0; ex += 1; }
        snprintf(ber, sizeof ber, "%.2fe%d", mant, ex);
    } else {
        snprintf(ber, sizeof ber, "0.00e0");
    }
    return snprintf(buf, buf_len, "%.2f %s %.1f %llu %llu %s +%.0f%% -%.0f%% 95%%", seconds, name ? name : "", ebn0_db,
                    (unsigned long long)frames, (unsigned long long)bit_errors, ber, up, dn);
}

// ------------------------------------------------------------------------------------------------------------------
// The tester's calling pattern (Utils.hs:53,62-69): maxThreadCount decoders, one frame per decode call, concurrent callers
// ------------------------------------------------------------------------------------------------------------------
extern "C" int ldpc_b200_ecc_tester_pattern(const char *codes_root, const char *name, int n_threads, int n_frames, const double *rx,
                                            int rx_len, uint8_t *out, double *seconds, uint64_t *mailbox_launches, uint64_t *mailbox_frames)
{
    if (!codes_root || !name || n_threads <= 0 || n_threads > 256 || n_frames < 0 || (n_frames && (!rx || !out)) || rx_len <= 0) {
        set_error("ecc_tester_pattern: bad arguments");
        return LDPC_B200_EINVAL;
    }
    std::vector<ldpc_b200_ecc_t *> recs(n_threads, nullptr);
    int rc = 0;
    for (int t = 0; t < n_threads && !rc; t++) rc = ldpc_b200_ecc_open(codes_root, name, 1, &recs[t]);
    const int K = rc ? 0 : recs[0]->K;
    // warm-up outside the clock: device copies of H, the mailbox, the kernels' first launch
    std::vector<uint8_t> scratch(K > 0 ? K : 1);
    for (int t = 0; t < n_threads && !rc && n_frames > 0; t++) rc = ldpc_b200_ecc_decode(recs[t], 1, rx, rx_len, scratch.data(), nullptr);
    uint64_t l0 = 0, f0 = 0;
    if (!rc && recs[0]->dec) ldpc_b200_decoder_mailbox_stats(recs[0]->dec, &l0, &f0);
    std::vector<int> rcs(n_threads, 0);
    std::vector<std::string> errs(n_threads);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    auto worker = [&](int t) {
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        for (int f = t; f < n_frames; f += n_threads) {
            int r = ldpc_b200_ecc_decode(recs[t], 1, rx + (size_t)f * rx_len, rx_len, out + (size_t)f * K, nullptr);
            if (r) { rcs[t] = r; errs[t] = ldpc_b200_last_error(); return; }
        }
    };
    double secs = 0.0;
    if (!rc) {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
        while (ready.load() < n_threads) std::this_thread::yield();
        const auto t0 = std::chrono::steady_clock::now();
        go.store(true, std::memory_order_release);
        for (auto &x : th) x.join();
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int t = 0; t < n_threads; t++) if (rcs[t]) { set_error("ecc_tester_pattern: thread %d: %s", t, errs[t].c_str()); rc = rcs[t]; break; }
    }
    if (!rc) {
        uint64_t l1 = 0, f1 = 0;
        if (recs[0]->dec) ldpc_b200_decoder_mailbox_stats(recs[0]->dec, &l1, &f1);
        if (mailbox_launches) *mailbox_launches = l1 - l0;
        if (mailbox_frames) *mailbox_frames = f1 - f0;
        if (seconds) *seconds = secs;
    }
    std::string keep = rc ? ldpc_b200_last_error() : "";
    for (auto *e : recs) if (e) ldpc_b200_ecc_close(e);
    if (rc) set_error("%s", keep.c_str());
    return rc;
}
