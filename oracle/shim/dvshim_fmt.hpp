// ORACLE shim (test infrastructure): stand-in for fmt::format as the reference's sources use it — "{}" placeholders replaced in
// order.  Integers print in decimal, floating-point values in the shortest form that reads back to the same value
// (std::to_chars; fmt's default is also a shortest round-trip form, its choice between fixed and exponent notation may differ),
// everything streamable through operator<<.  Format specs ("{:.2f}") are not interpreted: such call sites are log text only.
#pragma once
#include <charconv>
#include <sstream>
#include <string>
#include <type_traits>
namespace fmt {
namespace dvshim_detail {
template <class T>
inline std::string str(const T& v) {
    if constexpr (std::is_floating_point<T>::value) {
        char buf[64];
        auto r = std::to_chars(buf, buf + sizeof(buf), v);
        return std::string(buf, r.ptr);
    } else if constexpr (std::is_convertible<T, std::string>::value) {
        return std::string(v);
    } else {
        std::ostringstream os;
        os << v;
        return os.str();
    }
}
inline void fill(std::string&, size_t) {}
template <class T, class... A>
inline void fill(std::string& s, size_t from, const T& v, const A&... rest) {
    const size_t a = s.find('{', from);
    if (a == std::string::npos) return;
    const size_t b = s.find('}', a);
    if (b == std::string::npos) return;
    const std::string rep = str(v);
    s.replace(a, b - a + 1, rep);
    fill(s, a + rep.size(), rest...);
}
}  // namespace dvshim_detail
template <class... A>
inline std::string format(const std::string& f, const A&... args) {
    std::string s = f;
    dvshim_detail::fill(s, 0, args...);
    return s;
}
template <class... A>
inline std::string format(const char* f, const A&... args) { return format(std::string(f), args...); }
}  // namespace fmt
