// oracle/shim: the subset of boost::dynamic_bitset that libMems' match classes touch. Test infrastructure.
#pragma once
#include <vector>
#include <cstddef>
namespace boost {
template <class Block = unsigned long> class dynamic_bitset {
public:
	typedef size_t size_type;
	static const size_t npos = (size_t)-1;
	dynamic_bitset() {}
	explicit dynamic_bitset(size_t n, bool v = false) : bits(n, v) {}
	size_t size() const { return bits.size(); }
	void resize(size_t n, bool v = false) { bits.resize(n, v); }
	bool test(size_t i) const { return bits[i]; }
	bool operator[](size_t i) const { return bits[i]; }
	std::vector<bool>::reference operator[](size_t i) { return bits[i]; }
	dynamic_bitset& set(size_t i, bool v = true) { bits[i] = v; return *this; }
	dynamic_bitset& set() { bits.assign(bits.size(), true); return *this; }
	dynamic_bitset& reset(size_t i) { bits[i] = false; return *this; }
	dynamic_bitset& reset() { bits.assign(bits.size(), false); return *this; }
	dynamic_bitset& flip() { for (size_t i = 0; i < bits.size(); ++i) bits[i] = !bits[i]; return *this; }
	size_t count() const { size_t c = 0; for (size_t i = 0; i < bits.size(); ++i) c += bits[i]; return c; }
	bool any() const { return count() > 0; }
	size_t find_first() const { for (size_t i = 0; i < bits.size(); ++i) if (bits[i]) return i; return npos; }
	size_t find_next(size_t p) const { for (size_t i = p + 1; i < bits.size(); ++i) if (bits[i]) return i; return npos; }
private:
	std::vector<bool> bits;
};
}
