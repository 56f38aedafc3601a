// util::range — the integer / iterator ranges the reference's samples loop over
// (spice/include/spice/util/range.h:10-42): range(n) = [0, n), range(a, b) = [a, max(a, b)),
// range(first, last) over iterators, range(container) = [0, container.size()).
#pragma once

#include <algorithm>
#include <concepts>
#include <cstdint>
#include <iterator>

namespace spice::util {
template <class It>
struct range_t {
	It first, last;
	constexpr It begin() const { return first; }
	constexpr It end() const { return last; }
	constexpr std::int64_t size() const { return std::distance(first, last); }
};

// counts up; dereferences to the count
class int_iterator {
public:
	using iterator_category = std::random_access_iterator_tag;
	using iterator_concept  = std::random_access_iterator_tag;
	using difference_type   = std::int64_t;
	using value_type        = std::int64_t;
	using pointer           = std::int64_t const*;
	using reference         = std::int64_t;

	constexpr int_iterator() = default;
	constexpr int_iterator(std::int64_t at) : _at(at) {}
	constexpr std::int64_t operator*() const { return _at; }
	constexpr int_iterator& operator++() { return ++_at, *this; }
	constexpr int_iterator operator++(int) { return int_iterator(_at++); }
	constexpr int_iterator& operator--() { return --_at, *this; }
	constexpr int_iterator operator--(int) { return int_iterator(_at--); }
	constexpr int_iterator& operator+=(std::int64_t d) { return _at += d, *this; }
	constexpr int_iterator& operator-=(std::int64_t d) { return _at -= d, *this; }
	constexpr std::int64_t operator[](std::int64_t d) const { return _at + d; }
	friend constexpr int_iterator operator+(int_iterator a, std::int64_t d) { return int_iterator(a._at + d); }
	friend constexpr int_iterator operator+(std::int64_t d, int_iterator a) { return int_iterator(a._at + d); }
	friend constexpr int_iterator operator-(int_iterator a, std::int64_t d) { return int_iterator(a._at - d); }
	friend constexpr std::int64_t operator-(int_iterator a, int_iterator b) { return a._at - b._at; }
	friend constexpr bool operator==(int_iterator a, int_iterator b) { return a._at == b._at; }
	friend constexpr bool operator!=(int_iterator a, int_iterator b) { return a._at != b._at; }
	friend constexpr bool operator<(int_iterator a, int_iterator b) { return a._at < b._at; }
	friend constexpr bool operator>(int_iterator a, int_iterator b) { return a._at > b._at; }
	friend constexpr bool operator<=(int_iterator a, int_iterator b) { return a._at <= b._at; }
	friend constexpr bool operator>=(int_iterator a, int_iterator b) { return a._at >= b._at; }

private:
	std::int64_t _at = 0;
};

constexpr range_t<int_iterator> range(std::int64_t min, std::int64_t max) { return {int_iterator(min), int_iterator(std::max(min, max))}; }
constexpr range_t<int_iterator> range(std::int64_t max) { return range(0, max); }

template <std::input_iterator It>
constexpr auto range(It first, It last) {
	return range_t<It>{first, last};
}

template <class Container>
requires requires(Container c) {
	{ c.size() } -> std::integral;
}
constexpr auto range(Container const& c) {
	return range(static_cast<std::int64_t>(c.size()));
}
}
