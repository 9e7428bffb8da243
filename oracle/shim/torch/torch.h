// ORACLE shim (test infrastructure): stand-in for the few libtorch names the reference's front-end sources use.  The tracking path
// never touches a tensor; SemanticImage::SetMaskAndRoi / SetBackgroundMask (basic/semantic_image.cpp:20-117) do integer mask
// arithmetic on an N x H x W instance-mask tensor, which is implemented here for exactly those calls (to(kInt8 / kUInt8), abs,
// clamp, sum(0), * scalar, operator[], sizes, data_ptr: int8 / uint8 conversions wrap like the C++ casts libtorch performs).
// Everything else aborts when reached.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <memory>
#include <vector>
namespace torch {
[[noreturn]] inline void dvshim_no_tensor() { std::fprintf(stderr, "oracle/shim: this torch::Tensor operation is outside the parity path\n"); std::abort(); }
constexpr int kUInt8 = 0, kInt8 = 1, kInt32 = 3, kInt64 = 4, kFloat = 6;
struct IntArrayRef {
    std::vector<int64_t> v;
    IntArrayRef() {}
    IntArrayRef(std::initializer_list<int64_t> l) : v(l) {}
    IntArrayRef(const std::vector<int64_t>& l) : v(l) {}
    int64_t operator[](size_t i) const { return v[i]; }
    size_t size() const { return v.size(); }
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
    Tensor() {}
    // integer tensor of the given shape from row-major values (test harness)
    Tensor(const std::vector<int64_t>& shape, const int64_t* values, int dtype = kInt64) : shape_(shape), dtype_(dtype) {
        size_t n = 1;
        for (int64_t d : shape) n *= (size_t)d;
        val_ = std::make_shared<std::vector<int64_t>>(values, values + n);
        cast_in_place();
    }
    Tensor index(std::initializer_list<indexing::Slice>) const { dvshim_no_tensor(); }
    bool defined() const { return (bool)val_; }
    int64_t numel() const { return val_ ? (int64_t)val_->size() : 0; }
    Tensor sum() const { dvshim_no_tensor(); }
    Tensor sum(IntArrayRef) const { dvshim_no_tensor(); }
    Tensor sum(int64_t dim) const {                       // integer sums promote to int64
        if (!val_ || dim != 0 || shape_.size() < 2) dvshim_no_tensor();
        Tensor r;
        r.shape_.assign(shape_.begin() + 1, shape_.end());
        r.dtype_ = kInt64;
        const size_t inner = val_->size() / (size_t)shape_[0];
        r.val_ = std::make_shared<std::vector<int64_t>>(inner, 0);
        for (size_t i = 0; i < val_->size(); i++) (*r.val_)[i % inner] += (*val_)[i];
        return r;
    }
    Tensor abs() const { Tensor r = copy(); for (auto& x : *r.val_) x = x < 0 ? -x : x; r.cast_in_place(); return r; }
    Tensor clamp(int64_t lo, int64_t hi) const { Tensor r = copy(); for (auto& x : *r.val_) x = x < lo ? lo : (x > hi ? hi : x); return r; }
    Scalar item() const { dvshim_no_tensor(); }
    template <class T> T item() const { dvshim_no_tensor(); }
    Tensor operator*(const Tensor&) const { dvshim_no_tensor(); }
    Tensor operator*(int64_t k) const { Tensor r = copy(); for (auto& x : *r.val_) x *= k; r.cast_in_place(); return r; }
    Tensor operator[](int64_t i) const {
        if (!val_ || shape_.empty()) dvshim_no_tensor();
        Tensor r;
        r.shape_.assign(shape_.begin() + 1, shape_.end());
        r.dtype_ = dtype_;
        const size_t inner = val_->size() / (size_t)shape_[0];
        r.val_ = std::make_shared<std::vector<int64_t>>(val_->begin() + (size_t)i * inner, val_->begin() + (size_t)(i + 1) * inner);
        return r;
    }
    IntArrayRef sizes() const { return IntArrayRef(shape_); }
    int64_t size(int d) const { return shape_.at((size_t)d); }
    Tensor to(int dtype) const { Tensor r = copy(); r.dtype_ = dtype; r.cast_in_place(); return r; }
    Tensor clone() const { return copy(); }
    void* data_ptr() const {                              // contiguous bytes of a uint8 / int8 tensor
        if (!val_ || (dtype_ != kUInt8 && dtype_ != kInt8)) dvshim_no_tensor();
        bytes_ = std::make_shared<std::vector<uint8_t>>(val_->size());
        for (size_t i = 0; i < val_->size(); i++) (*bytes_)[i] = (uint8_t)(*val_)[i];
        return bytes_->data();
    }
private:
    Tensor copy() const {
        if (!val_) dvshim_no_tensor();
        Tensor r;
        r.shape_ = shape_; r.dtype_ = dtype_;
        r.val_ = std::make_shared<std::vector<int64_t>>(*val_);
        return r;
    }
    void cast_in_place() {                                // the value range of the dtype, wrapping like a C++ integer cast
        if (dtype_ == kInt8) for (auto& x : *val_) x = (int64_t)(int8_t)x;
        else if (dtype_ == kUInt8) for (auto& x : *val_) x = (int64_t)(uint8_t)x;
        else if (dtype_ == kInt32) for (auto& x : *val_) x = (int64_t)(int32_t)x;
    }
    std::vector<int64_t> shape_;
    int dtype_ = kInt64;
    std::shared_ptr<std::vector<int64_t>> val_;
    mutable std::shared_ptr<std::vector<uint8_t>> bytes_;
};
}
