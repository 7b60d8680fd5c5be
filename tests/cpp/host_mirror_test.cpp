// Exercises the C++ host mirror (phylogenetics_b200/host/phylo_b200.hpp) on the GPU with the
// RNG-free fixtures of the reference and closed forms (values: tests/golden/known_answers.json).
#include <cstdio>
#include <map>
#include <string>

#include "phylo_b200.hpp"

using namespace phylo_b200;

static int failures = 0;
static void expect_close(const char* what, double got, double want, double tol)
{
    const bool ok = std::fabs(got - want) <= tol * std::fabs(want);
    std::printf("%-58s got %.15f want %.15f %s\n", what, got, want, ok ? "ok" : "FAIL");
    if (!ok) failures++;
}

int main()
{
    namespace SEM = Site_evolution_model;
    // 1. JC69, two leaves at 0.3 and 0.7: closed form, through a `Mat;`Diag;`Mat decomposition
    {
        using T = Tree<int, int, double>;
        auto t = T::binary_node(0, T::branch(0.3, T::leaf(0)), T::branch(0.7, T::leaf(1)));
        const auto red = SEM::JC69::rate_matrix_reduction();
        std::function<Phylo_ctmc::matrix_decomposition(const double&)> tp = [&](const double& l) {
            Vec d(4);
            for (int i = 0; i < 4; i++) d[i] = std::exp(l * red.diag[i]);
            return Phylo_ctmc::matrix_decomposition{Phylo_ctmc::Op::mat(red.p), Phylo_ctmc::Op::diag(d),
                                                    Phylo_ctmc::Op::mat(red.p_inv)};
        };
        const Vec pi(4, 0.25);
        expect_close("Phylo_ctmc.pruning JC69 pair, same state",
                     Phylo_ctmc::pruning<int, int, double>(*t, 4, tp, [](const int&) { return 2; }, pi),
                     -2.189931069177981, 1e-14);
        expect_close("Phylo_ctmc.pruning JC69 pair, different states",
                     Phylo_ctmc::pruning<int, int, double>(*t, 4, tp, [](const int& l) { return l == 0 ? 0 : 3; }, pi),
                     -3.078566665552957, 1e-14);
    }
    // 2. lib/mCMC.ml:22,36-38: Felsenstein(K80) 2.0 on ((T0:0.1,T1:0.1):0.1,(T2:3.0,T3:0.1):0.1), columns A,A,A,T
    {
        using PT = Phylogenetic_tree;
        auto tree = PT::node(0.1, PT::node(0.1, PT::leaf("T0"), 0.1, PT::leaf("T1")), 0.1,
                             PT::node(3.0, PT::leaf("T2"), 0.1, PT::leaf("T3")));
        using Align = std::map<std::string, std::string>;
        Align align{{"T0", "A"}, {"T1", "A"}, {"T2", "A"}, {"T3", "T"}};
        Felsenstein<SEM::K80, Align> F;
        F.get_base = [](const Align& a, const std::string& seq, int64_t pos) { return (int)std::string("ACGT").find(a.at(seq)[pos]); };
        F.length = [](const Align& a) { return (int64_t)a.begin()->second.size(); };
        expect_close("Felsenstein(K80 kappa=2).felsenstein, mCMC fixture", F.felsenstein(2.0, *tree, align), -5.704499092002674, 1e-13);
        expect_close("Felsenstein(K80 kappa=2).felsenstein_noshift", F.felsenstein_noshift(2.0, *tree, align), -5.704499092002674, 1e-13);
    }
    // 3. Site_evolution_model.K80.transition_probability_matrix: rows sum to 1, P(0) = I
    {
        const auto ms = SEM::transition_probability_matrices<SEM::K80>(2.0, {0.0, 0.1});
        double rowsum = 0.0;
        for (int j = 0; j < 4; j++) rowsum += ms[1](2, j);
        expect_close("K80 P(0.1) row sum", rowsum, 1.0, 1e-15);
        expect_close("K80 P(0)(1,1)", ms[0](1, 1), 1.0, 1e-15);
    }
    // 4. error conventions: Invalid_argument for a bad state, Failure for a call-order error
    {
        using T = Tree<int, int, double>;
        auto t = T::binary_node(0, T::branch(0.3, T::leaf(0)), T::branch(0.7, T::leaf(1)));
        std::function<Phylo_ctmc::matrix_decomposition(const double&)> tp = [](const double&) {
            Mat id(4);
            for (int i = 0; i < 4; i++) id(i, i) = 1.0;
            return Phylo_ctmc::matrix_decomposition{Phylo_ctmc::Op::mat(id)};
        };
        bool threw = false;
        try { Phylo_ctmc::pruning<int, int, double>(*t, 4, tp, [](const int&) { return 7; }, Vec(4, 0.25)); }
        catch (const std::invalid_argument&) { threw = true; }
        std::printf("%-58s %s\n", "leaf state out of range -> invalid_argument", threw ? "ok" : "FAIL");
        failures += !threw;
        threw = false;
        try { Context c(4, 10, 2); c.loglik(); }
        catch (const Failure&) { threw = true; }
        std::printf("%-58s %s\n", "evaluation before inputs -> Failure", threw ? "ok" : "FAIL");
        failures += !threw;
    }
    // 5. 2-bit packed nucleotide upload == byte-per-site upload; both scale-factor rules; conditional simulation
    {
        using T = Tree<int, int, double>;
        auto t = T::binary_node(0, T::branch(0.4, T::binary_node(0, T::branch(0.8, T::leaf(0)), T::branch(1.2, T::leaf(1)))),
                                T::branch(0.1, T::binary_node(0, T::branch(0.6, T::leaf(2)), T::branch(0.3, T::leaf(3)))));
        std::vector<const int*> leaf_payloads;
        std::vector<const double*> branch_payloads;
        const FlatTree f = flatten(*t, leaf_payloads, branch_payloads);
        const int64_t nsites = 9;
        std::vector<uint8_t> tips(4 * nsites), packed(4 * 3, 0);
        for (int r = 0; r < 4; r++)
            for (int64_t s = 0; s < nsites; s++) {
                tips[r * nsites + s] = (uint8_t)((r * 7 + s * 3) & 3);
                packed[r * 3 + s / 4] |= (uint8_t)(tips[r * nsites + s] << (2 * (s & 3)));
            }
        const auto red = SEM::K80::rate_matrix_reduction(2.0);
        Vec bl((size_t)f.nnodes(), 0.0);
        for (int v = 0; v < f.nnodes(); v++) bl[v] = branch_payloads[v] ? *branch_payloads[v] : 0.0;
        double totals[3];
        std::vector<uint8_t> states;
        for (int k = 0; k < 3; k++) {
            Context c(4, nsites, 4, 1, 0, k == 2 ? PB_FLAG_KEEP_CLV : 0);
            c.set_tree(f, &bl);
            c.set_eigensystem(red.p, red.diag, red.p_inv);
            c.set_root_frequencies(Vec(4, 0.25));
            if (k == 1) { c.set_option("fused_pow2", 0); c.set_tips_packed2(packed, 3); }
            else c.set_tips(tips, nsites);
            totals[k] = c.loglik();
            if (k == 2) states = c.conditional_simulation(42);
        }
        expect_close("packed upload + 1/max factor == bytes + 2^-k factor", totals[1], totals[0], 1e-14);
        expect_close("level kernels (kept CLVs) == site-resident kernel", totals[2], totals[0], 1e-14);
        bool leaves_kept = true;
        for (int v = 0; v < f.nnodes(); v++)
            if (f.child_ptr[v] == f.child_ptr[v + 1])
                for (int64_t s = 0; s < nsites; s++) leaves_kept &= states[v * nsites + s] == tips[f.tip_row[v] * nsites + s];
        std::printf("%-58s %s\n", "conditional_simulation keeps the observed leaves", leaves_kept ? "ok" : "FAIL");
        failures += !leaves_kept;
    }
    // 6. Nelder_mead.minimize driving the GPU evaluator (the shape of BASELINE config 5 in miniature): the length of
    //    the branch above T2 of the mCMC fixture's tree, on a log scale, maximising Felsenstein(K80 kappa = 2)
    {
        using PT = Phylogenetic_tree;
        using Align = std::map<std::string, std::string>;
        Align align{{"T0", "AACGTTAC"}, {"T1", "AACGTTAC"}, {"T2", "AGCGTCAC"}, {"T3", "TACGATAC"}};
        Felsenstein<SEM::K80, Align> F;
        F.get_base = [](const Align& a, const std::string& seq, int64_t pos) { return (int)std::string("ACGT").find(a.at(seq)[pos]); };
        F.length = [](const Align& a) { return (int64_t)a.begin()->second.size(); };
        int evaluations = 0;
        auto objective = [&](const Nelder_mead::Point& x) {
            evaluations++;
            auto tree = PT::node(0.1, PT::node(0.1, PT::leaf("T0"), 0.1, PT::leaf("T1")), 0.1,
                                 PT::node(std::exp(x[0]), PT::leaf("T2"), 0.1, PT::leaf("T3")));
            return -F.felsenstein(2.0, *tree, align);
        };
        double start = -3.0;
        const auto r = Nelder_mead::minimize(objective, [&] { start += 1.0; return Nelder_mead::Point{start}; }, 1e-10, 200);
        const double f0 = objective(r.point);
        const bool stationary = objective({r.point[0] + 0.05}) >= f0 && objective({r.point[0] - 0.05}) >= f0;
        const bool improved = f0 <= objective({-2.0}) && f0 <= objective({-1.0});
        std::printf("%-58s t = %.6f, -logL = %.9f, %d evaluations %s\n", "Nelder_mead over one branch length on the GPU",
                    std::exp(r.point[0]), f0, evaluations, (stationary && improved) ? "ok" : "FAIL");
        failures += !(stationary && improved);
    }
    std::printf(failures ? "%d FAILED\n" : "all ok\n", failures);
    return failures ? 1 : 0;
}
