// oracle/shim: stand-in for libMUSCLE/threadstorage.h — the oracle is single-threaded, so the
// "thread-local" holder is a plain value. Test infrastructure.
#pragma once
template <class T> class TLS {
public:
	TLS() {}
	TLS(const T& t) : v(t) {}
	T& get() { return v; }
	const T& get() const { return v; }
private:
	T v;
};
