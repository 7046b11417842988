// PLY point-cloud files, in and out ("PLY in/out gives real-scan inputs").
// Host-only C++17, no dependency on the GPU library.  Drop-in for the point-cloud half of the
// reference's pcp/io/ply.hpp (read_ply :108-141 + :141-270, write_ply :279-457): same names,
// same template parameters, same return types, and the same bytes on disk —
//
//   ply / format {ascii|binary_little_endian|binary_big_endian} 1.0
//   element vertex N + property {float|double} x, y, z
//   element normal M + property {float|double} nx, ny, nz     (a SEPARATE element, after the vertices)
//   end_header, then N vertex records followed by M normal records;
//   ascii records are std::to_string(x) " " std::to_string(y) " " std::to_string(z) "\n".
//
// Differences, all on the permissive side: the binary payload moves in one read / write per
// element instead of one per point (a 100 M-point scan is 1.2 GB), `double` properties are
// honoured in binary files (the reference reads every binary component as a 4-byte float,
// io/ply.hpp:742-744), a header that ends without `end_header` or a truncated payload yields an
// empty cloud instead of garbage, and read_ply_flat() hands the coordinates over as the packed
// float rows the C ABI (pcpx.h) takes, without a per-point object.
#ifndef PCPX_PLY_HPP
#define PCPX_PLY_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <istream>
#include <ostream>
#include <sstream>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

namespace pcp {
namespace io {

enum class ply_format_t { ascii, binary_little_endian, binary_big_endian };
enum class ply_coordinate_type_t { single_precision, double_precision };

struct ply_parameters_t
{
    ply_format_t format                         = ply_format_t::ascii;
    std::size_t vertex_count                    = 0u;
    std::size_t normal_count                    = 0u;
    ply_coordinate_type_t vertex_component_type = ply_coordinate_type_t::single_precision;
    ply_coordinate_type_t normal_component_type = ply_coordinate_type_t::single_precision;
};

inline ply_format_t string_to_format(std::string const& s)
{
    if (s == "binary_little_endian")
        return ply_format_t::binary_little_endian;
    if (s == "binary_big_endian")
        return ply_format_t::binary_big_endian;
    return ply_format_t::ascii;
}

inline bool is_machine_little_endian()
{
    std::uint16_t const probe = 1u;
    unsigned char first;
    std::memcpy(&first, &probe, 1);
    return first == 1u;
}
inline bool is_machine_big_endian() { return !is_machine_little_endian(); }

namespace detail {

inline std::vector<std::string> words(std::string const& line)
{
    std::vector<std::string> out;
    std::istringstream ss(line);
    for (std::string w; ss >> w;)
        out.push_back(w);
    return out;
}

// the three `property <type> <name>` lines of an element; all of one type, names as given
inline bool read_triplet(std::istream& is, std::array<char const*, 3> const& names,
                         ply_coordinate_type_t& type)
{
    std::string line;
    for (int i = 0; i < 3; ++i)
    {
        if (!std::getline(is, line))
            return false;
        auto const w = words(line);
        if (w.size() < 3 || w[0] != "property" || w[1] == "list" || w.back() != names[i])
            return false;
        ply_coordinate_type_t const t = w[1] == "double" ? ply_coordinate_type_t::double_precision
                                                         : ply_coordinate_type_t::single_precision;
        if (i > 0 && t != type)
            return false;
        type = t;
    }
    return true;
}

// false = not a PLY point cloud this reader understands
inline bool read_header(std::istream& is, ply_parameters_t& prm)
{
    std::string line;
    if (!std::getline(is, line))
        return false;
    auto w = words(line);
    if (w.empty() || w[0] != "ply")
        return false;
    while (std::getline(is, line))
    {
        w = words(line);
        if (w.empty() || w[0] == "comment")
            continue;
        if (w[0] == "end_header")
            return true;
        if (w[0] == "format" && w.size() >= 2)
            prm.format = string_to_format(w[1]);
        else if (w[0] == "element" && w.size() >= 3 && w[1] == "vertex")
        {
            prm.vertex_count = std::stoull(w.back());
            if (!read_triplet(is, {"x", "y", "z"}, prm.vertex_component_type))
                return false;
        }
        else if (w[0] == "element" && w.size() >= 3 && w[1] == "normal")
        {
            prm.normal_count = std::stoull(w.back());
            if (!read_triplet(is, {"nx", "ny", "nz"}, prm.normal_component_type))
                return false;
        }
    }
    return false; // no end_header
}

template <class T>
T byteswapped(T v)
{
    unsigned char b[sizeof(T)];
    std::memcpy(b, &v, sizeof(T));
    for (std::size_t i = 0; i < sizeof(T) / 2; ++i)
        std::swap(b[i], b[sizeof(T) - 1 - i]);
    std::memcpy(&v, b, sizeof(T));
    return v;
}

// `count` records of three components -> packed float rows appended to out
inline bool read_records(std::istream& is, ply_parameters_t const& prm, std::size_t count,
                         ply_coordinate_type_t type, std::vector<float>& out)
{
    out.resize(3 * count);
    if (count == 0)
        return true;
    if (prm.format == ply_format_t::ascii)
    {
        std::string line;
        for (std::size_t i = 0; i < count; ++i)
        {
            if (!std::getline(is, line))
                return false;
            auto const w = words(line);
            if (w.size() < 3)
                return false;
            for (int c = 0; c < 3; ++c)
                out[3 * i + c] = std::stof(w[c]);
        }
        return true;
    }
    bool const swap = (prm.format == ply_format_t::binary_little_endian) != is_machine_little_endian();
    if (type == ply_coordinate_type_t::single_precision)
    {
        is.read(reinterpret_cast<char*>(out.data()), static_cast<std::streamsize>(12 * count));
        if (static_cast<std::size_t>(is.gcount()) != 12 * count)
            return false;
        if (swap)
            for (float& v : out)
                v = byteswapped(v);
        return true;
    }
    std::vector<double> wide(3 * count);
    is.read(reinterpret_cast<char*>(wide.data()), static_cast<std::streamsize>(24 * count));
    if (static_cast<std::size_t>(is.gcount()) != 24 * count)
        return false;
    for (std::size_t i = 0; i < wide.size(); ++i)
        out[i] = static_cast<float>(swap ? byteswapped(wide[i]) : wide[i]);
    return true;
}

} // namespace detail

// Coordinates (and normals, when the file has them) as packed float rows: 3 * N and 3 * M
// values.  Returns false, leaving both empty, when the stream is not a readable PLY cloud.
inline bool read_ply_flat(std::istream& is, std::vector<float>& xyz, std::vector<float>& normals,
                          ply_parameters_t* params = nullptr)
{
    ply_parameters_t prm;
    xyz.clear(), normals.clear();
    bool ok = detail::read_header(is, prm) &&
              detail::read_records(is, prm, prm.vertex_count, prm.vertex_component_type, xyz) &&
              detail::read_records(is, prm, prm.normal_count, prm.normal_component_type, normals);
    if (!ok)
        xyz.clear(), normals.clear();
    if (params)
        *params = prm;
    return ok;
}

inline bool read_ply_flat(std::filesystem::path const& path, std::vector<float>& xyz,
                          std::vector<float>& normals, ply_parameters_t* params = nullptr)
{
    xyz.clear(), normals.clear();
    if (!std::filesystem::exists(path) || path.extension() != ".ply")
        return false;
    std::ifstream fs(path.string(), std::ios::binary);
    return fs.is_open() && read_ply_flat(fs, xyz, normals, params);
}

// io/ply.hpp:141-270: vectors of Point / Normal objects; empty on any failure
template <class Point, class Normal>
inline auto read_ply(std::istream& is) -> std::tuple<std::vector<Point>, std::vector<Normal>>
{
    using pc = typename Point::coordinate_type;
    using nc = typename Normal::component_type;
    std::vector<float> xyz, nrm;
    if (!read_ply_flat(is, xyz, nrm))
        return {};
    std::vector<Point> vertices(xyz.size() / 3);
    std::vector<Normal> normals(nrm.size() / 3);
    for (std::size_t i = 0; i < vertices.size(); ++i)
        vertices[i] = Point{static_cast<pc>(xyz[3 * i]), static_cast<pc>(xyz[3 * i + 1]),
                            static_cast<pc>(xyz[3 * i + 2])};
    for (std::size_t i = 0; i < normals.size(); ++i)
        normals[i] = Normal{static_cast<nc>(nrm[3 * i]), static_cast<nc>(nrm[3 * i + 1]),
                            static_cast<nc>(nrm[3 * i + 2])};
    return std::make_tuple(std::move(vertices), std::move(normals));
}

// io/ply.hpp:108-126
template <class Point, class Normal>
inline auto read_ply(std::filesystem::path const& path)
    -> std::tuple<std::vector<Point>, std::vector<Normal>>
{
    if (!path.has_filename() || !path.has_extension() || path.extension() != ".ply" ||
        !std::filesystem::exists(path))
        return {};
    std::ifstream fs(path.string(), std::ios::binary);
    if (!fs.is_open())
        return {};
    return read_ply<Point, Normal>(fs);
}

namespace detail {

template <class T>
void write_records(std::ostream& os, ply_format_t format, std::vector<T> const& packed)
{
    if (format == ply_format_t::ascii)
    {
        std::string text;
        text.reserve(packed.size() * 10);
        for (std::size_t i = 0; i + 2 < packed.size(); i += 3)
        {
            text += std::to_string(packed[i]);
            text += ' ';
            text += std::to_string(packed[i + 1]);
            text += ' ';
            text += std::to_string(packed[i + 2]);
            text += '\n';
        }
        os << text;
        return;
    }
    bool const swap = (format == ply_format_t::binary_little_endian) != is_machine_little_endian();
    if (!swap)
    {
        os.write(reinterpret_cast<char const*>(packed.data()),
                 static_cast<std::streamsize>(packed.size() * sizeof(T)));
        return;
    }
    std::vector<T> other(packed.size());
    for (std::size_t i = 0; i < packed.size(); ++i)
        other[i] = byteswapped(packed[i]);
    os.write(reinterpret_cast<char const*>(other.data()),
             static_cast<std::streamsize>(other.size() * sizeof(T)));
}

inline char const* format_line(ply_format_t f)
{
    switch (f)
    {
    case ply_format_t::binary_little_endian: return "format binary_little_endian 1.0\n";
    case ply_format_t::binary_big_endian: return "format binary_big_endian 1.0\n";
    default: return "format ascii 1.0\n";
    }
}

} // namespace detail

// io/ply.hpp:311-457
template <class Point, class Normal>
inline void write_ply(std::ostream& os, std::vector<Point> const& vertices,
                      std::vector<Normal> const& normals, ply_format_t format = ply_format_t::ascii)
{
    using pc = typename Point::coordinate_type;
    using nc = typename Normal::component_type;
    char const* const pt = std::is_same_v<pc, double> ? "double" : "float";
    char const* const nt = std::is_same_v<nc, double> ? "double" : "float";
    os << "ply\n"
       << detail::format_line(format) << "element vertex " << vertices.size() << "\n"
       << "property " << pt << " x\nproperty " << pt << " y\nproperty " << pt << " z\n"
       << "element normal " << normals.size() << "\n"
       << "property " << nt << " nx\nproperty " << nt << " ny\nproperty " << nt << " nz\n"
       << "end_header\n";
    std::vector<pc> p(3 * vertices.size());
    for (std::size_t i = 0; i < vertices.size(); ++i)
        p[3 * i] = vertices[i].x(), p[3 * i + 1] = vertices[i].y(), p[3 * i + 2] = vertices[i].z();
    detail::write_records(os, format, p);
    std::vector<nc> q(3 * normals.size());
    for (std::size_t i = 0; i < normals.size(); ++i)
        q[3 * i] = normals[i].nx(), q[3 * i + 1] = normals[i].ny(), q[3 * i + 2] = normals[i].nz();
    detail::write_records(os, format, q);
}

// io/ply.hpp:279-298
template <class Point, class Normal>
inline void write_ply(std::filesystem::path const& path, std::vector<Point> const& vertices,
                      std::vector<Normal> const& normals, ply_format_t format = ply_format_t::ascii)
{
    if (!path.has_filename() || !path.has_extension() || path.extension() != ".ply")
        return;
    std::ofstream ofs(path.string(), std::ios::binary);
    if (ofs.is_open())
        write_ply<Point, Normal>(ofs, vertices, normals, format);
}

// packed float rows straight from / for the C ABI
inline void write_ply_flat(std::ostream& os, std::vector<float> const& xyz,
                           std::vector<float> const& normals,
                           ply_format_t format = ply_format_t::binary_little_endian)
{
    os << "ply\n"
       << detail::format_line(format) << "element vertex " << xyz.size() / 3 << "\n"
       << "property float x\nproperty float y\nproperty float z\n"
       << "element normal " << normals.size() / 3 << "\n"
       << "property float nx\nproperty float ny\nproperty float nz\n"
       << "end_header\n";
    detail::write_records(os, format, xyz);
    detail::write_records(os, format, normals);
}

} // namespace io
} // namespace pcp

#endif // PCPX_PLY_HPP
