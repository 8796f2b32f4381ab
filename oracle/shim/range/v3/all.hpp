// oracle/shim/range/v3/all.hpp — TEST INFRASTRUCTURE: eager stand-ins for the handful of
// range-v3 views/actions the reference uses.  view::sample / action::shuffle feed the RANSAC
// sampling only, which the recorded-list configurations replace.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <random>
#include <set>
#include <tuple>
#include <utility>
#include <vector>
namespace ranges {
struct range_access {};
struct default_sentinel {};
struct adaptor_base {};
template <typename D> struct view_facade {};
template <typename D, typename R> struct view_adaptor { view_adaptor() = default; template <typename X> view_adaptor(X&&) {} };
template <typename R> using iterator_t = decltype(std::begin(std::declval<R&>()));
struct to_vector_fn {};
static constexpr to_vector_fn to_vector{};
template <typename T> inline std::vector<T> operator|(std::vector<T> v, to_vector_fn) { return v; }
template <typename R> inline auto distance(const R& r) { return r.size(); }
template <typename R, typename I, typename F>
inline auto accumulate(R&& r, I init, F f) { auto acc = init; for (auto&& x : r) acc = f(acc, x); return acc; }
namespace view {
template <typename T> struct unbounded { T start; };
template <typename T> inline unbounded<T> ints(T a) { return {a}; }
template <typename T, typename U> inline std::vector<T> ints(T a, U b) { std::vector<T> v; for (T i = a; i < (T)b; ++i) v.push_back(i); return v; }
template <typename T, typename C>
inline auto zip(unbounded<T> u, C& c) {
    std::vector<std::pair<T, typename C::value_type&>> out;
    T i = u.start;
    for (auto& x : c) out.push_back({i++, x});
    return out;
}
template <typename C, typename F>
inline auto filter(const C& c, F f) { std::vector<typename C::value_type> out; for (auto& x : c) if (f(x)) out.push_back(x); return out; }
template <typename C, typename F>
inline auto transform(const C& c, F f) { std::vector<decltype(f(*c.begin()))> out; for (auto& x : c) out.push_back(f(x)); return out; }
template <typename C, typename G>
inline auto sample(const C& c, uint64_t n, G& g) {  // selection sampling, order preserving
    std::vector<typename C::value_type> out;
    uint64_t left = c.size();
    for (auto& x : c) {
        if (n == 0) break;
        if (std::uniform_int_distribution<uint64_t>(0, left - 1)(g) < n) { out.push_back(x); --n; }
        --left;
    }
    return out;
}
template <typename C>
inline auto chunk(const C& c, uint64_t n) {
    std::vector<std::vector<typename C::value_type>> out;
    for (size_t i = 0; i < c.size(); i += n) out.emplace_back(c.begin() + i, c.begin() + std::min<size_t>(c.size(), i + n));
    return out;
}
template <typename C>
inline auto tail(const C& c) { return std::vector<typename C::value_type>(c.empty() ? c.begin() : c.begin() + 1, c.end()); }
template <typename C>
inline auto take_exactly(const C& c, uint64_t n) { return std::vector<typename C::value_type>(c.begin(), c.begin() + n); }
template <typename A, typename B>
struct product_view {  // first range outermost, like range-v3's cartesian_product
    const A& a; const B& b;
    struct iterator {
        const product_view* p; size_t i, j;
        std::tuple<typename A::value_type, typename B::value_type> operator*() const { return {p->a[i], p->b[j]}; }
        iterator& operator++() { if (++j == p->b.size()) { j = 0; ++i; } return *this; }
        bool operator!=(const iterator& o) const { return i != o.i || j != o.j; }
    };
    iterator begin() const { return {this, (b.size() && a.size()) ? (size_t)0 : a.size(), 0}; }
    iterator end() const { return {this, a.size(), 0}; }
    size_t size() const { return a.size() * b.size(); }
};
template <typename A, typename B> inline product_view<A, B> cartesian_product(const A& a, const B& b) { return {a, b}; }
}  // namespace view
namespace action {
template <typename G> struct shuffle_fn { G& g; };
template <typename G> inline shuffle_fn<G> shuffle(G& g) { return {g}; }
template <typename C, typename G> inline C& operator|=(C& c, shuffle_fn<G> s) { std::shuffle(c.begin(), c.end(), s.g); return c; }
}  // namespace action
}  // namespace ranges
