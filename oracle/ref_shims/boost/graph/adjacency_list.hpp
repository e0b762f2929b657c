// oracle/ref_shims/boost/graph/adjacency_list.hpp — TEST INFRASTRUCTURE ONLY.
//
// Container-only stand-in for the part of Boost.Graph that /root/reference/src/Stereo3DMST.cpp
// uses (it calls no Boost algorithm): adjacency_list<vecS|setS, vecS, undirectedS, VP, EP> with
// add_edge / edge / adjacent_vertices / num_vertices / bundled-property operator[].
// Semantics kept: vecS out-edge lists iterate in insertion order; setS out-edge sets are unique
// and iterate by ascending target; add_edge grows the vertex set to max(u,v)+1 (vecS vertices).
#pragma once
#include <algorithm>
#include <tuple>
#include <utility>
#include <vector>

namespace boost {
struct vecS {};
struct setS {};
struct undirectedS {};
struct no_property {};
using std::tie;

template <class OutEdgeS, class VertexS, class Dir, class VP = no_property, class EP = no_property>
class adjacency_list {
public:
    typedef int vertex_descriptor;
    struct edge_descriptor {
        int u, v, idx;
    };
    struct out_edge {
        int target, idx;
    };
    struct adjacency_iterator {
        const out_edge* p;
        int operator*() const { return p->target; }
        adjacency_iterator& operator++() {
            ++p;
            return *this;
        }
        bool operator!=(const adjacency_iterator& o) const { return p != o.p; }
        bool operator==(const adjacency_iterator& o) const { return p == o.p; }
    };
    adjacency_list() {}
    explicit adjacency_list(size_t n) : vprops(n), out(n) {}
    VP& operator[](int v) { return vprops[v]; }
    EP& operator[](const edge_descriptor& e) { return eprops[e.idx]; }
    std::vector<VP> vprops;
    std::vector<std::vector<out_edge>> out;
    std::vector<EP> eprops;
    static constexpr bool unique_sorted = std::is_same<OutEdgeS, setS>::value;
};

template <class G> struct graph_traits {
    typedef typename G::adjacency_iterator adjacency_iterator;
    typedef typename G::vertex_descriptor vertex_descriptor;
    typedef typename G::edge_descriptor edge_descriptor;
};

template <class O, class V, class D, class VP, class EP>
size_t num_vertices(const adjacency_list<O, V, D, VP, EP>& g) {
    return g.out.size();
}

template <class O, class V, class D, class VP, class EP>
std::pair<typename adjacency_list<O, V, D, VP, EP>::edge_descriptor, bool> add_edge(int u, int v,
                                                                                    adjacency_list<O, V, D, VP, EP>& g) {
    typedef adjacency_list<O, V, D, VP, EP> G;
    const size_t need = (size_t)std::max(u, v) + 1;
    if (g.out.size() < need) {
        g.out.resize(need);
        g.vprops.resize(need);
    }
    if (G::unique_sorted) {
        for (const auto& oe : g.out[u])
            if (oe.target == v) return {typename G::edge_descriptor{u, v, oe.idx}, false};
    }
    const int idx = (int)g.eprops.size();
    g.eprops.emplace_back();
    auto ins = [&](int a, int b) {
        auto& lst = g.out[a];
        if (G::unique_sorted) {
            auto it = std::lower_bound(lst.begin(), lst.end(), b,
                                       [](const typename G::out_edge& x, int t) { return x.target < t; });
            lst.insert(it, typename G::out_edge{b, idx});
        } else
            lst.push_back(typename G::out_edge{b, idx});
    };
    ins(u, v);
    if (u != v) ins(v, u);
    return {typename G::edge_descriptor{u, v, idx}, true};
}

template <class O, class V, class D, class VP, class EP>
std::pair<typename adjacency_list<O, V, D, VP, EP>::edge_descriptor, bool> edge(int u, int v,
                                                                                adjacency_list<O, V, D, VP, EP>& g) {
    typedef adjacency_list<O, V, D, VP, EP> G;
    for (const auto& oe : g.out[u])
        if (oe.target == v) return {typename G::edge_descriptor{u, v, oe.idx}, true};
    return {typename G::edge_descriptor{u, v, -1}, false};
}

template <class O, class V, class D, class VP, class EP>
std::pair<typename adjacency_list<O, V, D, VP, EP>::adjacency_iterator,
          typename adjacency_list<O, V, D, VP, EP>::adjacency_iterator>
adjacent_vertices(int v, const adjacency_list<O, V, D, VP, EP>& g) {
    typedef typename adjacency_list<O, V, D, VP, EP>::adjacency_iterator It;
    if ((size_t)v >= g.out.size()) return {It{nullptr}, It{nullptr}};  // isolated tree: Boost would be UB here
    const auto& lst = g.out[v];
    return {It{lst.data()}, It{lst.data() + lst.size()}};
}
}  // namespace boost
