// TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product.
//
// Minimal stand-in for the three Boost.Iostreams headers that the reference's
// dxyWindow.cpp includes (/root/reference/dxyWindow.cpp:17-19).  Boost headers
// are not installed in this image, so oracle/Makefile compiles the UNMODIFIED
// reference source with `-I oracle/shim`.  Only the surface dxyWindow.cpp
// touches is provided (dxyWindow.cpp:256-278):
//   boost::iostreams::input
//   boost::iostreams::gzip_decompressor            (default constructible)
//   boost::iostreams::filtering_streambuf<input>   (push(filter), push(istream&))
// The gzip path is real: underflow() inflates through zlib so `.mafs.gz`
// inputs behave as with genuine Boost.
#ifndef PGT_ORACLE_SHIM_FILTERING_STREAMBUF_HPP
#define PGT_ORACLE_SHIM_FILTERING_STREAMBUF_HPP

#include <cstring>   // the reference calls strcmp without including it (dxyWindow.cpp:99)
#include <istream>
#include <streambuf>
#include <vector>
#include <zlib.h>

namespace boost {
namespace iostreams {

struct input {};

struct gzip_decompressor {};

template <class Mode>
class filtering_streambuf : public std::streambuf {
public:
	filtering_streambuf() : src_(0), zinit_(false), zeof_(false), ibuf_(1 << 16), obuf_(1 << 16) {
		std::memset(&zs_, 0, sizeof(zs_));
	}
	~filtering_streambuf() {
		if (zinit_) inflateEnd(&zs_);
	}
	void push(const gzip_decompressor&) {}
	void push(std::istream& is) {
		src_ = &is;
		// 15 window bits + 32 = auto-detect zlib/gzip header
		zinit_ = (inflateInit2(&zs_, 15 + 32) == Z_OK);
		setg(&obuf_[0], &obuf_[0], &obuf_[0]);
	}

protected:
	int_type underflow() {
		if (gptr() < egptr()) return traits_type::to_int_type(*gptr());
		if (!src_ || !zinit_ || zeof_) return traits_type::eof();
		for (;;) {
			if (zs_.avail_in == 0) {
				src_->read(&ibuf_[0], ibuf_.size());
				std::streamsize got = src_->gcount();
				if (got <= 0) return traits_type::eof();
				zs_.next_in = reinterpret_cast<Bytef*>(&ibuf_[0]);
				zs_.avail_in = static_cast<uInt>(got);
			}
			zs_.next_out = reinterpret_cast<Bytef*>(&obuf_[0]);
			zs_.avail_out = static_cast<uInt>(obuf_.size());
			int rc = inflate(&zs_, Z_NO_FLUSH);
			size_t produced = obuf_.size() - zs_.avail_out;
			if (rc == Z_STREAM_END) {
				// concatenated gzip members (bgzip output) continue with a fresh header
				if (zs_.avail_in > 0 || src_->peek() != traits_type::eof()) inflateReset(&zs_);
				else zeof_ = true;
			} else if (rc != Z_OK && rc != Z_BUF_ERROR) {
				zeof_ = true;
			}
			if (produced > 0) {
				setg(&obuf_[0], &obuf_[0], &obuf_[0] + produced);
				return traits_type::to_int_type(*gptr());
			}
			if (zeof_) return traits_type::eof();
		}
	}

private:
	std::istream* src_;
	z_stream zs_;
	bool zinit_;
	bool zeof_;
	std::vector<char> ibuf_;
	std::vector<char> obuf_;
};

}  // namespace iostreams
}  // namespace boost

#endif
