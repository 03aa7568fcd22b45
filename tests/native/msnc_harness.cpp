// TEST INFRASTRUCTURE ONLY: runs phynetpy_b200/csrc/pnb_msnc_core.hpp -- the network compiler, the
// gene-tree preparation and the very DP source the CUDA kernel K-m executes -- on the host, so the
// CPU test-suite can check them against the oracle and the golden vectors without a GPU.
// Built by tests/test_msnc_host.py with g++; never part of libphynet_b200.so.
#include "../../phynetpy_b200/csrc/pnb_msnc_core.hpp"

#include <cstdio>

using namespace pnb::msnc;

extern "C" int msnc_host_eval(int n_nodes, int n_edges, const int32_t* edge_src, const int32_t* edge_dst,
                              const double* edge_gamma, const double* height, const int32_t* node_species,
                              int n_species, int64_t n_trees, const int64_t* tree_off, const int32_t* parent,
                              const double* blen, const int32_t* gene_species, double theta, int cap, int force_nw,
                              int stride_in, const double* edge_theta, double root_theta, int tree_path,
                              double* out, int32_t* info /* W, n_steps, max status, NW used */, char* err, int err_cap) {
    NetProgram prog;
    std::string e = compile_network(n_nodes, n_edges, edge_src, edge_dst, edge_gamma, height, node_species, n_species, prog);
    if (!e.empty()) {
        snprintf(err, err_cap, "%s", e.c_str());
        return 1;
    }
    int max_n = 0;
    for (int64_t j = 0; j < n_trees; ++j) max_n = std::max<int>(max_n, static_cast<int>(tree_off[j + 1] - tree_off[j]));
    const int NW = force_nw ? force_nw : (max_n <= 64 ? 1 : 2);
    info[0] = prog.W;
    info[1] = static_cast<int32_t>(prog.steps.size());
    info[2] = 0;
    info[3] = NW;
    std::vector<double> step_theta(4 * prog.steps.size());
    fill_step_theta(prog, theta, edge_theta, root_theta, step_theta.data());
    TreeScratch sc;
    std::vector<double> ev_t(std::max(max_n, 1));
    std::vector<uint32_t> ev_n(std::max(max_n, 1));
    std::vector<uint64_t> sp_mask(static_cast<size_t>(std::max(n_species, 1)) * NW);
    // stride > 1 emulates the interleaved shared-memory layout of the kernel (element k at [k * stride])
    const size_t stride = stride_in > 0 ? static_cast<size_t>(stride_in) : 1;
    std::vector<uint64_t> fr_mask(msnc_mask_words(prog.W, NW, cap) * stride);
    std::vector<double> fr_lp(msnc_lp_words(cap) * stride);
    for (int64_t j = 0; j < n_trees; ++j) {
        const int n = static_cast<int>(tree_off[j + 1] - tree_off[j]);
        if (n == 0) {
            out[j] = 0.0;   // empty gene tree, src/_msnc_density.py:222-223
            continue;
        }
        int n_ev = 0, root_bit = 0;
        const char* te = prepare_tree(n, parent + tree_off[j], blen + tree_off[j], gene_species + tree_off[j], n_species, NW, sc,
                                      ev_t.data(), ev_n.data(), &n_ev, sp_mask.data(), &root_bit);
        if (te) {
            snprintf(err, err_cap, "tree %lld: %s", static_cast<long long>(j), te);
            return 1;
        }
        int status = 0;
        if (tree_path && prog.retic_step.empty()) {   // the single-entry fast path the kernel takes for species trees
            out[j] = NW == 1
                ? msnc_dp_tree<1>(prog.steps.data(), static_cast<int>(prog.steps.size()), prog.down_slot.data(), ev_t.data(), ev_n.data(),
                                  n_ev, sp_mask.data(), root_bit, step_theta.data(), fr_mask.data(), stride)
                : msnc_dp_tree<2>(prog.steps.data(), static_cast<int>(prog.steps.size()), prog.down_slot.data(), ev_t.data(), ev_n.data(),
                                  n_ev, sp_mask.data(), root_bit, step_theta.data(), fr_mask.data(), stride);
            continue;
        }
        if (NW == 1)
            out[j] = msnc_dp<1>(prog.steps.data(), static_cast<int>(prog.steps.size()), prog.down_slot.data(), prog.W, cap,
                                ev_t.data(), ev_n.data(), n_ev, sp_mask.data(), root_bit, step_theta.data(), fr_mask.data(),
                                fr_lp.data(), stride, &status);
        else
            out[j] = msnc_dp<2>(prog.steps.data(), static_cast<int>(prog.steps.size()), prog.down_slot.data(), prog.W, cap,
                                ev_t.data(), ev_n.data(), n_ev, sp_mask.data(), root_bit, step_theta.data(), fr_mask.data(),
                                fr_lp.data(), stride, &status);
        info[2] = std::max(info[2], status);
    }
    return 0;
}

extern "C" double msnc_host_frontier_bound(int n_nodes, int n_edges, const int32_t* edge_src, const int32_t* edge_dst,
                                           const double* edge_gamma, const double* height, const int32_t* node_species,
                                           int n_species, const int32_t* alleles) {
    NetProgram prog;
    if (!compile_network(n_nodes, n_edges, edge_src, edge_dst, edge_gamma, height, node_species, n_species, prog).empty())
        return -1.0;
    return frontier_bound(prog, alleles);
}
