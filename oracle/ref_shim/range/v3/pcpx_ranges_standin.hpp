// Minimal stand-in for the three range-v3 facilities the reference's hot-path headers
// use (views::zip of two std::array, views::transform, make_subrange). TEST
// INFRASTRUCTURE ONLY: it exists so that oracle/ref_bridge.cpp can compile the
// UNMODIFIED reference headers (range-v3 0.11.0 is fetched by the reference's CMake and is
// not vendored under /root/reference). Written from the usage sites, not from range-v3:
//   common/axis_aligned_bounding_box.hpp:48,64   zip + std::all_of
//   common/norm.hpp:132-135                      zip | transform, std::copy
//   kdtree/linked_kdtree.hpp:462                 zip + std::all_of
//   octree/linked_octree.hpp:114                 make_subrange | transform -> bounding_box
#ifndef PCPX_RANGES_STANDIN_HPP
#define PCPX_RANGES_STANDIN_HPP

#include <cstddef>
#include <iterator>
#include <tuple>
#include <type_traits>
#include <utility>

namespace ranges {

template <class It>
class subrange_standin
{
  public:
    subrange_standin(It b, It e) : b_(b), e_(e) {}
    It begin() const { return b_; }
    It end() const { return e_; }

  private:
    It b_, e_;
};

template <class It>
subrange_standin<It> make_subrange(It b, It e)
{
    return subrange_standin<It>(b, e);
}

namespace views {

// ---- zip -------------------------------------------------------------------------------
template <class A, class B>
class zip_view_standin
{
    using ia = decltype(std::begin(std::declval<A const&>()));
    using ib = decltype(std::begin(std::declval<B const&>()));

  public:
    class iterator
    {
      public:
        using iterator_category = std::forward_iterator_tag;
        using value_type =
            std::tuple<typename std::iterator_traits<ia>::value_type,
                       typename std::iterator_traits<ib>::value_type>;
        using difference_type = std::ptrdiff_t;
        using pointer         = value_type const*;
        using reference       = value_type;

        iterator() = default;
        iterator(ia a, ib b) : a_(a), b_(b) {}
        reference operator*() const { return value_type(*a_, *b_); }
        iterator& operator++()
        {
            ++a_;
            ++b_;
            return *this;
        }
        iterator operator++(int)
        {
            iterator old = *this;
            ++*this;
            return old;
        }
        bool operator==(iterator const& o) const { return a_ == o.a_; }
        bool operator!=(iterator const& o) const { return a_ != o.a_; }

      private:
        ia a_{};
        ib b_{};
    };

    zip_view_standin(A const& a, B const& b) : a_(&a), b_(&b) {}
    iterator begin() const { return iterator(std::begin(*a_), std::begin(*b_)); }
    iterator end() const { return iterator(std::end(*a_), std::end(*b_)); }

  private:
    A const* a_;
    B const* b_;
};

template <class A, class B>
zip_view_standin<A, B> zip(A const& a, B const& b)
{
    return zip_view_standin<A, B>(a, b);
}

// ---- transform -------------------------------------------------------------------------
template <class F>
struct transform_closure_standin
{
    F f;
};

template <class F>
transform_closure_standin<std::decay_t<F>> transform(F&& f)
{
    return {std::forward<F>(f)};
}

template <class Rng, class F>
class transform_view_standin
{
    using base_iter = decltype(std::declval<Rng const&>().begin());

  public:
    class iterator
    {
      public:
        using iterator_category = std::forward_iterator_tag;
        using reference  = decltype(std::declval<F const&>()(*std::declval<base_iter const&>()));
        using value_type = std::decay_t<reference>;
        using difference_type = std::ptrdiff_t;
        using pointer         = value_type const*;

        iterator() = default;
        iterator(base_iter it, F const* f) : it_(it), f_(f) {}
        reference operator*() const { return (*f_)(*it_); }
        iterator& operator++()
        {
            ++it_;
            return *this;
        }
        iterator operator++(int)
        {
            iterator old = *this;
            ++*this;
            return old;
        }
        bool operator==(iterator const& o) const { return it_ == o.it_; }
        bool operator!=(iterator const& o) const { return it_ != o.it_; }

      private:
        base_iter it_{};
        F const* f_ = nullptr;
    };

    transform_view_standin(Rng rng, F f) : rng_(std::move(rng)), f_(std::move(f)) {}
    iterator begin() const { return iterator(rng_.begin(), &f_); }
    iterator end() const { return iterator(rng_.end(), &f_); }

  private:
    Rng rng_;
    F f_;
};

template <class Rng, class F>
transform_view_standin<std::decay_t<Rng>, F> operator|(Rng&& rng, transform_closure_standin<F> c)
{
    return transform_view_standin<std::decay_t<Rng>, F>(std::forward<Rng>(rng), std::move(c.f));
}

} // namespace views
} // namespace ranges

#endif // PCPX_RANGES_STANDIN_HPP
