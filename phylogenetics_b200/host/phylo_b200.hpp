// C++ host mirror of the reference's interface for the likelihood path, on top of the C ABI
// (include/phylo_b200.h).  The reference is compiled code (OCaml) whose toolchain is absent from
// this image, so the compiled-language host layer is C++17: same names, argument meaning and error
// behaviour as lib/phylo_ctmc.mli, lib/felsenstein.mli, lib/site_evolution_model.mli.
// Header-only; link against libphylo_b200.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "phylo_b200.h"

namespace phylo_b200 {

// PB_ERR_INVALID_ARG -> std::invalid_argument (OCaml Invalid_argument); the rest -> Failure.
struct Failure : std::runtime_error { using std::runtime_error::runtime_error; };

inline void check(pb_ctx* c, int rc)
{
    if (rc == PB_OK) return;
    const std::string msg = pb_last_error(c);
    if (rc == PB_ERR_INVALID_ARG) throw std::invalid_argument(msg);
    throw Failure(msg);
}

// Column-major n x n matrix, the layout of the reference's Lacaml mat (lib/linear_algebra.ml:151-152).
struct Mat {
    int n = 0;
    std::vector<double> a;
    Mat() = default;
    explicit Mat(int n_) : n(n_), a((size_t)n_ * n_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)j * n + i]; }
    double operator()(int i, int j) const { return a[(size_t)j * n + i]; }
    static Mat of_rows(const std::vector<std::vector<double>>& r)
    {
        Mat m((int)r.size());
        for (int i = 0; i < m.n; i++)
            for (int j = 0; j < m.n; j++) m(i, j) = r[i][j];
        return m;
    }
};
using Vec = std::vector<double>;

inline Mat dot(const Mat& x, const Mat& y)       // Matrix.dot = dgemm
{
    Mat r(x.n);
    for (int i = 0; i < x.n; i++)
        for (int j = 0; j < x.n; j++) {
            double acc = 0.0;
            for (int k = 0; k < x.n; k++) acc += x(i, k) * y(k, j);
            r(i, j) = acc;
        }
    return r;
}

// ---- Tree.t (lib/tree.ml:3-13): k-ary, payloads on nodes, leaves and branches -------------
template <class N, class L, class B>
struct Tree {
    struct Branch { B data; std::shared_ptr<Tree> tip; };
    bool is_leaf = true;
    L leaf_data{};
    N node_data{};
    std::vector<Branch> branches;    // List1: non-empty for a node

    static std::shared_ptr<Tree> leaf(L l)
    {
        auto t = std::make_shared<Tree>();
        t->leaf_data = std::move(l);
        return t;
    }
    static std::shared_ptr<Tree> node(N d, std::vector<Branch> bs)
    {
        if (bs.empty()) throw std::invalid_argument("Tree.node: empty branch list");
        auto t = std::make_shared<Tree>();
        t->is_leaf = false;
        t->node_data = std::move(d);
        t->branches = std::move(bs);
        return t;
    }
    static std::shared_ptr<Tree> binary_node(N d, Branch l, Branch r) { return node(std::move(d), {std::move(l), std::move(r)}); }
    static Branch branch(B d, std::shared_ptr<Tree> tip) { return Branch{std::move(d), std::move(tip)}; }
};

struct FlatTree {
    int root = 0;
    std::vector<int32_t> child_ptr{0}, child_idx, tip_row;
    int nnodes() const { return (int)tip_row.size(); }
};

// Pre-order numbering; leaves[row] and branches[v] (payload of the branch above node v) by pointer.
template <class N, class L, class B>
FlatTree flatten(const Tree<N, L, B>& t, std::vector<const L*>& leaves, std::vector<const B*>& branches)
{
    if (t.is_leaf) throw std::invalid_argument("the root must be an internal node");
    FlatTree f;
    struct Item { const Tree<N, L, B>* t; const B* b; };
    std::vector<Item> order;
    std::vector<std::vector<int>> kids;
    std::function<int(const Tree<N, L, B>&, const B*)> visit = [&](const Tree<N, L, B>& x, const B* b) {
        const int id = (int)order.size();
        order.push_back({&x, b});
        kids.emplace_back();
        if (!x.is_leaf)
            for (const auto& br : x.branches) { const int c = visit(*br.tip, &br.data); kids[id].push_back(c); }
        return id;
    };
    visit(t, nullptr);
    for (size_t v = 0; v < order.size(); v++) {
        branches.push_back(order[v].b);
        if (order[v].t->is_leaf) { f.tip_row.push_back((int32_t)leaves.size()); leaves.push_back(&order[v].t->leaf_data); }
        else f.tip_row.push_back(-1);
        for (int c : kids[v]) f.child_idx.push_back(c);
        f.child_ptr.push_back((int32_t)f.child_idx.size());
    }
    return f;
}

// ---- RAII context over the C ABI --------------------------------------------------------
class Context {
public:
    Context(int nstates, int64_t nsites, int ntaxa, int ncat = 1, int device = 0, unsigned flags = 0)
        : n_(nstates), nsites_(nsites)
    {
        c_ = pb_create(device, nstates, ncat, nsites, ntaxa, flags);
        if (!c_) {
            const std::string msg = pb_last_error(nullptr);
            if (msg.find("must be") != std::string::npos) throw std::invalid_argument(msg);
            throw Failure(msg);
        }
    }
    ~Context() { pb_destroy(c_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    pb_ctx* raw() { return c_; }
    void set_mode(int m) { check(c_, pb_set_mode(c_, m)); }
    void set_rescaling(double thr, bool root) { check(c_, pb_set_rescaling(c_, thr, root)); }
    void set_tree(const FlatTree& f, const Vec* bl)
    {
        check(c_, pb_set_tree(c_, f.nnodes(), f.root, f.child_ptr.data(), f.child_idx.data(), f.tip_row.data(),
                              bl ? bl->data() : nullptr));
        nnodes_ = f.nnodes();
    }
    void set_tips(const std::vector<uint8_t>& tips, int64_t ld) { check(c_, pb_set_tips(c_, tips.data(), ld)); }
    // nucleotides 2-bit packed, 4 sites per byte (pb_set_tips_packed2)
    void set_tips_packed2(const std::vector<uint8_t>& packed, int64_t ld) { check(c_, pb_set_tips_packed2(c_, packed.data(), ld)); }
    void set_option(const char* name, double value) { check(c_, pb_set_option(c_, name, value)); }
    // Phylo_ctmc.conditional_simulation for every site: states[node * nsites + site] (needs PB_FLAG_KEEP_CLV)
    std::vector<uint8_t> conditional_simulation(uint64_t seed)
    {
        std::vector<uint8_t> states((size_t)nnodes_ * (size_t)nsites_);
        check(c_, pb_conditional_simulation(c_, seed, states.data(), nsites_));
        return states;
    }
    void set_root_frequencies(const Vec& pi) { check(c_, pb_set_root_frequencies(c_, pi.data())); }
    void set_eigensystem(const Mat& u, const Vec& lambda, const Mat& uinv)
    {
        check(c_, pb_set_eigensystem(c_, u.a.data(), lambda.data(), uinv.a.data()));
    }
    void set_pmatrices(const std::vector<double>& p) { check(c_, pb_set_pmatrices(c_, p.data())); }
    double loglik(std::vector<double>* per_site = nullptr)
    {
        double total = 0.0;
        if (per_site) per_site->resize((size_t)nsites_);
        check(c_, pb_loglik(c_, &total, per_site ? per_site->data() : nullptr));
        return total;
    }

private:
    pb_ctx* c_ = nullptr;
    int n_, nnodes_ = 0;
    int64_t nsites_;
};

// ---- Phylo_ctmc (lib/phylo_ctmc.mli) -----------------------------------------------------
namespace Phylo_ctmc {

struct Op {                       // one factor of a matrix_decomposition (lib/phylo_ctmc.ml:4-8)
    enum Kind { MatK, TransposeK, DiagK } kind;
    Mat m;
    Vec d;
    static Op mat(Mat x) { return Op{MatK, std::move(x), {}}; }
    static Op transpose(Mat x) { return Op{TransposeK, std::move(x), {}}; }
    static Op diag(Vec v) { return Op{DiagK, Mat(), std::move(v)}; }
};
using matrix_decomposition = std::vector<Op>;

// lib/phylo_ctmc.ml:13-24
inline Mat matrix_decomposition_reduce(int dim, const matrix_decomposition& d, size_t from = 0)
{
    auto dense = [&](const Op& o) {
        if (o.kind == Op::MatK) return o.m;
        Mat r(dim);
        if (o.kind == Op::TransposeK) {
            for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) r(i, j) = o.m(j, i);
        } else {
            for (int i = 0; i < dim; i++) r(i, i) = o.d[i];
        }
        return r;
    };
    if (from >= d.size()) { Mat id(dim); for (int i = 0; i < dim; i++) id(i, i) = 1.0; return id; }
    if (from + 1 == d.size()) return dense(d[from]);
    return dot(dense(d[from]), matrix_decomposition_reduce(dim, d, from + 1));
}

// Batched form: total and per-site log-likelihoods; leaf_state(leaf, site).
template <class N, class L, class B>
double pruning_sites(const Tree<N, L, B>& t, int nstates, int64_t nsites,
                     const std::function<matrix_decomposition(const B&)>& transition_probabilities,
                     const std::function<int(const L&, int64_t)>& leaf_state, const Vec& root_frequencies,
                     std::vector<double>* per_site = nullptr)
{
    std::vector<const L*> leaves;
    std::vector<const B*> branches;
    const FlatTree f = flatten(t, leaves, branches);
    const int nn = f.nnodes(), n = nstates;
    Context ctx(n, nsites, (int)leaves.size());
    ctx.set_tree(f, nullptr);
    std::vector<double> P((size_t)nn * n * n, 0.0);
    for (int v = 0; v < nn; v++) {
        if (!branches[v]) continue;
        const Mat m = matrix_decomposition_reduce(n, transition_probabilities(*branches[v]));
        if (m.n != n) throw std::invalid_argument("transition matrix has the wrong shape");
        std::copy(m.a.begin(), m.a.end(), P.begin() + (size_t)v * n * n);
    }
    ctx.set_pmatrices(P);
    std::vector<uint8_t> tips(leaves.size() * (size_t)nsites);
    for (size_t r = 0; r < leaves.size(); r++)
        for (int64_t s = 0; s < nsites; s++) {
            const int st = leaf_state(*leaves[r], s);
            if (st < 0 || st >= n) throw std::invalid_argument("leaf state out of range");
            tips[r * nsites + s] = (uint8_t)st;
        }
    ctx.set_tips(tips, nsites);
    ctx.set_mode(PB_MODE_PHYLO_CTMC);
    ctx.set_root_frequencies(root_frequencies);
    return ctx.loglik(per_site);
}

// Phylo_ctmc.pruning (lib/phylo_ctmc.mli:55-61): one site.
template <class N, class L, class B>
double pruning(const Tree<N, L, B>& t, int nstates,
               const std::function<matrix_decomposition(const B&)>& transition_probabilities,
               const std::function<int(const L&)>& leaf_state, const Vec& root_frequencies)
{
    return pruning_sites<N, L, B>(t, nstates, 1, transition_probabilities,
                                  [&](const L& l, int64_t) { return leaf_state(l); }, root_frequencies);
}

}  // namespace Phylo_ctmc

// ---- Site_evolution_model (lib/site_evolution_model.ml) ----------------------------------
namespace Site_evolution_model {

struct Reduction { Mat p; Vec diag; Mat p_inv; };

struct JC69 {                      // :53-86
    using param = std::monostate;
    static Reduction rate_matrix_reduction(param = {})
    {
        return {Mat::of_rows({{1, -1, -1, -1}, {1, 1, 0, 0}, {1, 0, 1, 0}, {1, 0, 0, 1}}),
                {0.0, -4.0 / 3.0, -4.0 / 3.0, -4.0 / 3.0},
                Mat::of_rows({{0.25, 0.25, 0.25, 0.25}, {-0.25, 0.75, -0.25, -0.25}, {-0.25, -0.25, 0.75, -0.25},
                              {-0.25, -0.25, -0.25, 0.75}})};
    }
    static Vec stationary_distribution(param = {}) { return Vec(4, 0.25); }
};

struct K80 {                       // :93-130
    using param = double;
    static Reduction rate_matrix_reduction(param k)
    {
        const double l3 = (-2.0 * k - 2.0) / (k + 2.0);
        return {Mat::of_rows({{1, -1, 0, -1}, {1, 1, -1, 0}, {1, -1, 0, 1}, {1, 1, 1, 0}}),
                {0.0, -4.0 / (k + 2.0), l3, l3},
                Mat::of_rows({{0.25, 0.25, 0.25, 0.25}, {-0.25, 0.25, -0.25, 0.25}, {0, -0.5, 0, 0.5}, {-0.5, 0, 0.5, 0}})};
    }
    static Vec stationary_distribution(param) { return Vec(4, 0.25); }
};

// Make_diag.transition_probability_matrix (:43-47), one or many branch lengths, on the device (kernel K1).
template <class M>
std::vector<Mat> transition_probability_matrices(typename M::param p, const Vec& ts, int device = 0)
{
    const Reduction r = M::rate_matrix_reduction(p);
    const int n = r.p.n;
    std::vector<double> out(ts.size() * (size_t)n * n);
    check(nullptr, pb_transition_matrices_eigen(device, n, (int)ts.size(), r.p.a.data(), r.diag.data(),
                                                r.p_inv.a.data(), ts.data(), out.data()));
    std::vector<Mat> res;
    for (size_t k = 0; k < ts.size(); k++) {
        Mat m(n);
        std::copy(out.begin() + k * n * n, out.begin() + (k + 1) * n * n, m.a.begin());
        res.push_back(std::move(m));
    }
    return res;
}

}  // namespace Site_evolution_model

// ---- Felsenstein (lib/felsenstein.mli): binary Phylogenetic_tree, string-indexed alignment -------
struct Phylogenetic_tree {         // lib/phylogenetic_tree.ml:6-10
    bool is_leaf = true;
    std::string index;
    double l1 = 0, l2 = 0;
    std::shared_ptr<Phylogenetic_tree> left, right;
    static std::shared_ptr<Phylogenetic_tree> leaf(std::string i)
    {
        auto t = std::make_shared<Phylogenetic_tree>();
        t->index = std::move(i);
        return t;
    }
    static std::shared_ptr<Phylogenetic_tree> node(double l1, std::shared_ptr<Phylogenetic_tree> a, double l2,
                                                   std::shared_ptr<Phylogenetic_tree> b)
    {
        auto t = std::make_shared<Phylogenetic_tree>();
        t->is_leaf = false;
        t->l1 = l1; t->l2 = l2; t->left = std::move(a); t->right = std::move(b);
        return t;
    }
};

// Felsenstein.Make(A)(Align)(E): get_base(align, seq, pos) -> state, length(align).
template <class Model, class Align>
struct Felsenstein {
    std::function<int(const Align&, const std::string&, int64_t)> get_base;
    std::function<int64_t(const Align&)> length;

    double evaluate(typename Model::param p, const Phylogenetic_tree& tree, const Align& align, double threshold) const
    {
        using T = Tree<int, std::string, double>;
        std::function<std::shared_ptr<T>(const Phylogenetic_tree&)> conv = [&](const Phylogenetic_tree& t) {
            if (t.is_leaf) return T::leaf(t.index);
            return T::binary_node(0, T::branch(t.l1, conv(*t.left)), T::branch(t.l2, conv(*t.right)));
        };
        const auto tt = conv(tree);
        std::vector<const std::string*> leaves;
        std::vector<const double*> branches;
        const FlatTree f = flatten(*tt, leaves, branches);
        const int64_t nsites = length(align);
        Vec bl(f.nnodes(), 0.0);
        for (int v = 0; v < f.nnodes(); v++) if (branches[v]) bl[v] = *branches[v];
        Context ctx(4, nsites, (int)leaves.size());
        ctx.set_tree(f, &bl);
        const auto r = Model::rate_matrix_reduction(p);
        ctx.set_eigensystem(r.p, r.diag, r.p_inv);
        std::vector<uint8_t> tips(leaves.size() * (size_t)nsites);
        for (size_t row = 0; row < leaves.size(); row++)
            for (int64_t s = 0; s < nsites; s++) tips[row * nsites + s] = (uint8_t)get_base(align, *leaves[row], s);
        ctx.set_tips(tips, nsites);
        ctx.set_rescaling(threshold, false);           // shift_normal, no shift after x pi (lib/felsenstein.ml:47-48,57-64)
        ctx.set_root_frequencies(Model::stationary_distribution(p));
        return ctx.loglik();
    }
    double felsenstein(typename Model::param p, const Phylogenetic_tree& t, const Align& a) const { return evaluate(p, t, a, 0.0000001); }
    double felsenstein_noshift(typename Model::param p, const Phylogenetic_tree& t, const Align& a) const { return evaluate(p, t, a, -INFINITY); }
};

// Nelder_mead.minimize (lib/nelder_mead.ml:53-132): the derivative-free simplex search that turns one likelihood
// evaluation into a workload (BASELINE config 5).  alpha = 1, gamma = 2, rho = sigma = 0.5; the search stops when the
// sample standard deviation of the simplex values (Gsl.Stats.sd) falls below `tol` or after `maxit` iterations.
// `f` is opaque: with the GPU evaluator it uploads O(branches) scalars and reads one double back per call.
namespace Nelder_mead {

using Point = std::vector<double>;

inline Point centroid(const std::vector<Point>& xs)                       // :40-47
{
    if (xs.empty()) throw std::invalid_argument("Nelder_mead.centroid: empty array");
    Point c(xs[0].size());
    for (size_t i = 0; i < c.size(); i++) {
        double acc = 0.0;
        for (const Point& x : xs) acc = acc + x[i];
        c[i] = acc / (double)xs.size();
    }
    return c;
}

inline Point update(const Point& from, double alpha, const Point& towards)  // :49-51
{
    Point y(from.size());
    for (size_t i = 0; i < y.size(); i++) y[i] = from[i] + alpha * (towards[i] - from[i]);
    return y;
}

inline double sd(const std::vector<double>& v)         // Gsl.Stats.sd: running mean and variance, n - 1 denominator
{
    double mean = 0.0, var = 0.0;
    for (size_t i = 0; i < v.size(); i++) mean += (v[i] - mean) / (double)(i + 1);
    for (size_t i = 0; i < v.size(); i++) { const double d = v[i] - mean; var += (d * d - var) / (double)(i + 1); }
    return std::sqrt(var * ((double)v.size() / ((double)v.size() - 1.0)));
}

inline std::vector<int> array_order(const std::vector<double>& v)         // Utils.array_order: ranks, ties by index
{
    std::vector<int> r(v.size());
    for (size_t i = 0; i < r.size(); i++) r[i] = (int)i;
    std::stable_sort(r.begin(), r.end(), [&](int a, int b) { return v[a] < v[b]; });
    return r;
}

struct Result { double value; Point point; int iterations; };

template <class F, class S>
Result minimize(F&& f, S&& sample, double tol = 1e-8, int maxit = 100000)
{
    const double alpha = 1.0, gamma = 2.0, rho = 0.5, sigma = 0.5;
    const Point first = sample();          // like the reference, the first draw only fixes the dimension
    const size_t n = first.size();
    if (n == 0) throw std::invalid_argument("Nelder_mead.minimize: sample returns empty vectors");
    std::vector<Point> pts(n + 1);
    std::vector<double> obj(n + 1);
    for (Point& p : pts) {
        p = sample();
        if (p.size() != n) throw std::invalid_argument("Nelder_mead.minimize: sample returns vectors of varying lengths");
    }
    for (size_t k = 0; k <= n; k++) obj[k] = f(pts[k]);
    for (int it = 0;; it++) {
        const std::vector<int> rk = array_order(obj);
        const int best = rk[0], worst = rk[n];
        const double f_second_worst = obj[rk[n - 1]];
        std::vector<Point> keep;
        for (size_t k = 0; k < n; k++) keep.push_back(pts[rk[k]]);
        const Point c = centroid(keep);
        const Point x_r = update(c, -alpha, pts[worst]);
        const double f_r = f(x_r);
        if (f_r < obj[best]) {                                             // expansion
            const Point x_e = update(c, gamma, x_r);
            const double f_e = f(x_e);
            pts[worst] = f_e < f_r ? x_e : x_r;
            obj[worst] = std::fmin(f_r, f_e);
        } else if (f_r < f_second_worst) {                                 // reflection
            pts[worst] = x_r;
            obj[worst] = f_r;
        } else {
            const bool outside = f_r < obj[worst];
            const Point x_c = update(c, rho, outside ? x_r : pts[worst]);
            const double f_c = f(x_c);
            if (outside ? f_c <= f_r : f_c < obj[worst]) {                 // contraction
                pts[worst] = x_c;
                obj[worst] = f_c;
            } else {                                                       // shrink towards the best point
                for (size_t k = 0; k <= n; k++) {
                    if ((int)k == best) continue;
                    pts[k] = update(pts[best], sigma, pts[k]);
                    obj[k] = f(pts[k]);
                }
            }
        }
        if (sd(obj) < tol || it >= maxit) return Result{obj[best], pts[best], it};
    }
}

}  // namespace Nelder_mead

}  // namespace phylo_b200
