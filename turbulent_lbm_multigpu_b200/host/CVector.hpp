/*
 * CVector.hpp -- the small fixed-size vectors the solver surface is written in.
 *
 * Stand-in for the reference's src/libmath/CVector{3,4}.hpp restricted to what the hot path
 * uses: CVector<3,int>, CVector<3,T>, CVector<4,T>, operator[],
 * elements(), dotProd(), length(), scaling by a scalar, stream output "[a, b, c]".
 * One generic template instead of the reference's per-dimension classes.
 */
#ifndef LBM_B200_HOST_CVECTOR_HPP
#define LBM_B200_HOST_CVECTOR_HPP

#include <cmath>
#include <ostream>

template <int N, typename T>
class CVector {
public:
	T data[N];

	CVector() { for (int i = 0; i < N; i++) data[i] = T(0); }
	CVector(T a, T b, T c) { static_assert(N == 3, "three components"); data[0] = a; data[1] = b; data[2] = c; }
	CVector(T a, T b, T c, T d) { static_assert(N == 4, "four components"); data[0] = a; data[1] = b; data[2] = c; data[3] = d; }
	explicit CVector(const T *p) { for (int i = 0; i < N; i++) data[i] = p[i]; }
	template <typename U>
	explicit CVector(const CVector<N, U> &o) { for (int i = 0; i < N; i++) data[i] = (T)o.data[i]; }

	T &operator[](int i) { return data[i]; }
	const T &operator[](int i) const { return data[i]; }

	/* product of the components: the cell count of a size vector (CVector3.hpp:177-180) */
	T elements() const { T p = data[0]; for (int i = 1; i < N; i++) p *= data[i]; return p; }
	T dotProd(const CVector &o) const { T s = data[0] * o.data[0]; for (int i = 1; i < N; i++) s += data[i] * o.data[i]; return s; }
	/* evaluated in T, left to right, like the reference (CVector3.hpp:185-188) */
	T length() const { return (T)std::sqrt(dotProd(*this)); }
	T max() const { T m = data[0]; for (int i = 1; i < N; i++) if (data[i] > m) m = data[i]; return m; }

	CVector operator*(T s) const { CVector r; for (int i = 0; i < N; i++) r.data[i] = data[i] * s; return r; }
	CVector operator+(const CVector &o) const { CVector r; for (int i = 0; i < N; i++) r.data[i] = data[i] + o.data[i]; return r; }
	CVector operator-(const CVector &o) const { CVector r; for (int i = 0; i < N; i++) r.data[i] = data[i] - o.data[i]; return r; }
	bool operator==(const CVector &o) const { for (int i = 0; i < N; i++) if (!(data[i] == o.data[i])) return false; return true; }
	bool operator!=(const CVector &o) const { return !(*this == o); }
};

template <int N, typename T>
std::ostream &operator<<(std::ostream &os, const CVector<N, T> &v)
{
	os << "[";
	for (int i = 0; i < N; i++) os << (i ? ", " : "") << v[i];
	return os << "]";
}

#endif
