// Command-line driver of include/pcpx/ply.hpp for the pytest suite (host only, no GPU):
//   ply_tool read  <in.ply>  <out.bin>        flat read: u64 N, u64 M, N*3 f32, M*3 f32
//   ply_tool write <in.bin>  <out.ply> <fmt>  typed write_ply<point_t, normal_t>, fmt = 0|1|2
//   ply_tool copy  <in.ply>  <out.ply> <fmt>  typed read_ply then write_ply (round trip)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>

#include <pcpx/ply.hpp>

namespace {
struct point_t
{
    using coordinate_type = float;
    float x_ = 0, y_ = 0, z_ = 0;
    point_t() = default;
    point_t(float x, float y, float z) : x_(x), y_(y), z_(z) {}
    float x() const { return x_; }
    float y() const { return y_; }
    float z() const { return z_; }
};
struct normal_t
{
    using component_type = float;
    float x_ = 0, y_ = 0, z_ = 0;
    normal_t() = default;
    normal_t(float x, float y, float z) : x_(x), y_(y), z_(z) {}
    float nx() const { return x_; }
    float ny() const { return y_; }
    float nz() const { return z_; }
};
pcp::io::ply_format_t fmt_of(char const* s)
{
    return s[0] == '0'   ? pcp::io::ply_format_t::ascii
           : s[0] == '1' ? pcp::io::ply_format_t::binary_little_endian
                         : pcp::io::ply_format_t::binary_big_endian;
}
} // namespace

int main(int argc, char** argv)
{
    if (argc < 4)
        return 2;
    if (!std::strcmp(argv[1], "read"))
    {
        std::vector<float> xyz, nrm;
        bool const ok = pcp::io::read_ply_flat(std::filesystem::path(argv[2]), xyz, nrm);
        std::ofstream out(argv[3], std::ios::binary);
        std::uint64_t const n = xyz.size() / 3, m = nrm.size() / 3;
        out.write(reinterpret_cast<char const*>(&n), 8);
        out.write(reinterpret_cast<char const*>(&m), 8);
        out.write(reinterpret_cast<char const*>(xyz.data()), (std::streamsize)(xyz.size() * 4));
        out.write(reinterpret_cast<char const*>(nrm.data()), (std::streamsize)(nrm.size() * 4));
        return ok ? 0 : 1;
    }
    if (!std::strcmp(argv[1], "write") && argc >= 5)
    {
        std::ifstream in(argv[2], std::ios::binary);
        std::uint64_t n = 0, m = 0;
        in.read(reinterpret_cast<char*>(&n), 8);
        in.read(reinterpret_cast<char*>(&m), 8);
        std::vector<float> xyz(3 * n), nrm(3 * m);
        in.read(reinterpret_cast<char*>(xyz.data()), (std::streamsize)(xyz.size() * 4));
        in.read(reinterpret_cast<char*>(nrm.data()), (std::streamsize)(nrm.size() * 4));
        std::vector<point_t> p(n);
        std::vector<normal_t> q(m);
        for (std::uint64_t i = 0; i < n; ++i)
            p[i] = point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        for (std::uint64_t i = 0; i < m; ++i)
            q[i] = normal_t{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
        pcp::io::write_ply<point_t, normal_t>(std::filesystem::path(argv[3]), p, q, fmt_of(argv[4]));
        return 0;
    }
    if (!std::strcmp(argv[1], "copy") && argc >= 5)
    {
        auto [p, q] = pcp::io::read_ply<point_t, normal_t>(std::filesystem::path(argv[2]));
        pcp::io::write_ply<point_t, normal_t>(std::filesystem::path(argv[3]), p, q, fmt_of(argv[4]));
        return 0;
    }
    return 2;
}
