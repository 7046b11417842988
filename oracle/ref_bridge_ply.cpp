// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product path.
//
// Flat C bridge over the UNMODIFIED reference PLY reader / writer, compiled where the header
// lies under /root/reference/include by oracle/Makefile into oracle/_ref/libpcp_ref_ply.so:
//   pcp::io::write_ply<Point, Normal>(std::ostream&, ...)   include/pcp/io/ply.hpp:311-457
//   pcp::io::read_ply<Point, Normal>(std::istream&)         include/pcp/io/ply.hpp:141-270
// Used to produce the golden byte streams of tests/golden/ref_ply.npz and, when present, as
// the live cross-check of include/pcpx/ply.hpp.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include <pcp/common/normals/normal.hpp>
#include <pcp/common/points/point.hpp>
#include <pcp/io/ply.hpp>

extern "C" {

// format: 0 ascii, 1 binary little endian, 2 binary big endian.  Returns the byte count; copies
// at most `capacity` bytes into out.
std::size_t ref_write_ply(const float* xyz, std::size_t n, const float* nrm, std::size_t m,
                          int format, unsigned char* out, std::size_t capacity)
{
    std::vector<pcp::point_t> points(n);
    std::vector<pcp::normal_t> normals(m);
    for (std::size_t i = 0; i < n; ++i)
        points[i] = pcp::point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    for (std::size_t i = 0; i < m; ++i)
        normals[i] = pcp::normal_t{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    std::ostringstream os(std::ios::binary);
    pcp::io::ply_format_t const f = format == 0   ? pcp::io::ply_format_t::ascii
                                    : format == 1 ? pcp::io::ply_format_t::binary_little_endian
                                                  : pcp::io::ply_format_t::binary_big_endian;
    pcp::io::write_ply<pcp::point_t, pcp::normal_t>(os, points, normals, f);
    std::string const s = os.str();
    std::memcpy(out, s.data(), std::min(capacity, s.size()));
    return s.size();
}

// counts[0] = vertices, counts[1] = normals read; rows copied up to the given capacities
void ref_read_ply(const unsigned char* bytes, std::size_t len, float* xyz, std::size_t cap_points,
                  float* nrm, std::size_t cap_normals, std::size_t* counts)
{
    std::istringstream is(std::string(reinterpret_cast<const char*>(bytes), len), std::ios::binary);
    auto [points, normals] = pcp::io::read_ply<pcp::point_t, pcp::normal_t>(is);
    counts[0] = points.size(), counts[1] = normals.size();
    for (std::size_t i = 0; i < std::min(points.size(), cap_points); ++i)
        xyz[3 * i] = points[i].x(), xyz[3 * i + 1] = points[i].y(), xyz[3 * i + 2] = points[i].z();
    for (std::size_t i = 0; i < std::min(normals.size(), cap_normals); ++i)
        nrm[3 * i] = normals[i].nx(), nrm[3 * i + 1] = normals[i].ny(),
                nrm[3 * i + 2] = normals[i].nz();
}

} // extern "C"
