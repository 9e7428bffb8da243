// ORACLE shim (test infrastructure): the front-end's parity path never touches a tensor; the type and the few members
// the reference headers name only have to exist (they abort when reached)
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <vector>
namespace torch {
[[noreturn]] inline void dvshim_no_tensor() { std::fprintf(stderr, "oracle/shim: torch::Tensor is outside the parity path\n"); std::abort(); }
struct IntArrayRef {
    IntArrayRef() {}
    IntArrayRef(std::initializer_list<int64_t>) {}
    IntArrayRef(const std::vector<int64_t>&) {}
};
struct Scalar {
    float toFloat() const { dvshim_no_tensor(); }
    double toDouble() const { dvshim_no_tensor(); }
    int toInt() const { dvshim_no_tensor(); }
    int64_t toLong() const { dvshim_no_tensor(); }
};
namespace indexing {
struct NoneType {};
static const NoneType None{};
struct Slice {
    Slice() {}
    template <class A> Slice(const A&) {}
    template <class A, class B> Slice(const A&, const B&) {}
};
}
class Tensor {
public:
    Tensor index(std::initializer_list<indexing::Slice>) const { dvshim_no_tensor(); }
    bool defined() const { return false; }
    int64_t numel() const { return 0; }
    Tensor sum() const { dvshim_no_tensor(); }
    Tensor sum(IntArrayRef) const { dvshim_no_tensor(); }
    Scalar item() const { dvshim_no_tensor(); }
    template <class T> T item() const { dvshim_no_tensor(); }
    Tensor operator*(const Tensor&) const { dvshim_no_tensor(); }
    Tensor operator[](int64_t) const { dvshim_no_tensor(); }
    IntArrayRef sizes() const { dvshim_no_tensor(); }
    int64_t size(int) const { dvshim_no_tensor(); }
    Tensor to(int) const { dvshim_no_tensor(); }
    Tensor clone() const { dvshim_no_tensor(); }
    void* data_ptr() const { dvshim_no_tensor(); }
};
}
