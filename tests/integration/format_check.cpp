// format_check.cpp -- TEST: the CLIs' number formatting (pgt_cli.h put_g / put_u32 / put_i32) against what the
// reference prints, `std::cout << double` with default settings (fstWindow.cpp:88, dxyWindow.cpp:190) == printf("%g").
// Values: random bit patterns, random decimals of 1..17 digits over 40 decades, exact rounding ties at the 6th
// significant digit, powers of ten and their neighbours, subnormals, zeros, infinities, NaN.
#include <cmath>
#include <cstring>
#include <random>
#include <sstream>

#include "pgt_cli.h"

static unsigned long long g_checked = 0;

static int check(double v) {
	char a[64], b[64];
	char* e = pgtcli::put_g(a, v);
	*e = 0;
	snprintf(b, sizeof(b), "%g", v);
	++g_checked;
	if (strcmp(a, b) != 0) {
		// printf prints "-nan" for a NaN with the sign bit set; the ostream does the same -- anything else is a bug
		fprintf(stderr, "MISMATCH value %.17g (bits %016llx): put_g '%s' printf '%s'\n", v, (unsigned long long)*(uint64_t*)&v, a, b);
		return 1;
	}
	return 0;
}

int main(int argc, char** argv) {
	const unsigned long long n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000ull;
	std::mt19937_64 rng(12345);
	int bad = 0;
	for (unsigned long long i = 0; i < n && bad < 10; ++i) {
		uint64_t bits = rng();
		double v;
		memcpy(&v, &bits, 8);
		bad += check(v);
		// decimal with d significant digits at a random decade
		const int d = 1 + (int)(rng() % 17);
		unsigned long long m = rng() % 100000000000000000ull;
		for (int k = 17; k > d; --k) m /= 10;
		const int ex = (int)(rng() % 40) - 20;
		bad += check((double)m * std::pow(10.0, ex));
		bad += check(-(double)m * std::pow(10.0, ex));
		// a tie at the 6th significant digit: ddddd5 exactly (binary-exact for small exponents), and its neighbours
		const double tie = (double)(100000 + rng() % 900000) + 0.5;
		for (int s = -3; s <= 3; ++s) {
			const double t = std::ldexp(tie, s * 3);
			bad += check(t);
			bad += check(std::nextafter(t, 0.0));
			bad += check(std::nextafter(t, 1e300));
		}
	}
	// put_g's fast path (|v| in [1e-5, 1e15): six digits from one scaled multiply / divide) hands every value within
	// 1e-6 of a rounding boundary to the general routine: sweep both sides of that margin, and the decade edges, in every
	// decade it covers
	{
		const double off[] = {0.0, 0.25, 0.49, 0.4999, 0.499998, 0.4999989, 0.4999991, 0.5, 0.5000009, 0.5000011, 0.500002, 0.5001, 0.51, 0.75, 0.999999};
		for (unsigned long long i = 0; i < n / 8 + 1000 && bad < 10; ++i) {
			const uint32_t six = i % 7 == 0 ? 999999u : (i % 7 == 1 ? 100000u : 100000u + (uint32_t)(rng() % 900000));
			for (int X = -6; X <= 15; ++X) {
				for (double o : off) {
					const long double scaled = (long double)six + (long double)o;
					const double v = (double)(X >= 5 ? scaled * powl(10.0L, X - 5) : scaled / powl(10.0L, 5 - X));
					bad += check(v);
					bad += check(-v);
					bad += check(std::nextafter(v, 0.0));
					bad += check(std::nextafter(v, 1e300));
				}
			}
		}
	}
	for (int ex = -320; ex <= 308; ++ex) {
		const double p = std::pow(10.0, ex);
		bad += check(p);
		bad += check(std::nextafter(p, 0.0));
		bad += check(std::nextafter(p, 1e300));
		bad += check(9.999995 * p);
		bad += check(9.9999949999 * p);
		bad += check(0.0001 * p);
	}
	const double special[] = {0.0, -0.0, 1.0, -1.0, 0.5, 100000.0, 999999.0, 999999.5, 1000000.0, 0.0001, 0.00009999995, 0.357143, 1e-05,
	                          2798.1, HUGE_VAL, -HUGE_VAL, 4.9406564584124654e-324, 2.2250738585072014e-308, 1.7976931348623157e308};
	for (double v : special) bad += check(v);
	bad += check(std::nan(""));
	// one value through the ostream itself, to pin "default ostream == %g" on this libstdc++
	std::ostringstream os;
	os << 0.357142857142857 << ' ' << 1e-05 << ' ' << 123456789.0 << ' ' << 0.1 + 0.2;
	if (os.str() != "0.357143 1e-05 1.23457e+08 0.3") {
		fprintf(stderr, "ostream default formatting is not %%g here: '%s'\n", os.str().c_str());
		++bad;
	}
	// integers
	char t[32];
	*pgtcli::put_u32(t, 4294967295u) = 0;
	bad += strcmp(t, "4294967295") != 0;
	*pgtcli::put_i32(t, -2147483647 - 1) = 0;
	bad += strcmp(t, "-2147483648") != 0;
	*pgtcli::put_u32(t, 0u) = 0;
	bad += strcmp(t, "0") != 0;
	printf("%llu values checked, %d mismatches\n", g_checked, bad);
	return bad ? 1 : 0;
}
