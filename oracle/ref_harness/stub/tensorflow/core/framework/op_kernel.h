// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Minimal stand-in for the two TensorFlow headers that the reference's
// external/structural_losses/tf_nndistance.cpp includes, so that the reference
// translation unit can be compiled UNMODIFIED, straight from /root/reference,
// without TensorFlow.  It models only what that file touches: Tensor (a view
// over caller memory), TensorShape, OpKernel / OpKernelContext, the
// OP_REQUIRES* macros and the REGISTER_* macros (which become no-ops).
//
// Written for this repo; it contains no TensorFlow code.
#pragma once
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>

namespace tensorflow {

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<long long> d) : d_(d) {}
  int dims() const { return (int)d_.size(); }
  long long dim_size(int i) const { return d_[i]; }
  long long num_elements() const {
    long long n = 1;
    for (long long v : d_) n *= v;
    return n;
  }
  bool operator==(const TensorShape& o) const { return d_ == o.d_; }

 private:
  std::vector<long long> d_;
};

template <typename T>
struct FlatView {
  T* p;
  T& operator()(long long i) const { return p[i]; }
};

// A non-owning view: the harness points it at the caller's buffers.
class Tensor {
 public:
  Tensor() : data_(nullptr) {}
  Tensor(void* data, const TensorShape& s) : data_(data), shape_(s) {}
  int dims() const { return shape_.dims(); }
  const TensorShape& shape() const { return shape_; }
  template <typename T>
  FlatView<T> flat() const { return FlatView<T>{static_cast<T*>(data_)}; }

 private:
  void* data_;
  TensorShape shape_;
};

struct Status {
  bool ok_ = true;
  std::string msg_;
  bool ok() const { return ok_; }
  static Status OK() { return Status(); }
};

namespace errors {
inline Status InvalidArgument(const char* m) {
  Status s;
  s.ok_ = false;
  s.msg_ = m;
  return s;
}
}  // namespace errors

class OpKernelConstruction {};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  // outputs[i] must be pre-pointed at caller memory; allocate_output just
  // re-labels it with the shape the op asks for and checks the capacity.
  std::vector<void*> out_ptr;
  std::vector<long long> out_capacity;  // elements
  Tensor outputs[8];  // fixed storage: the op keeps Tensor* across allocate_output calls
  Status status;

  const Tensor& input(int i) const { return inputs[i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** t) {
    if (i >= 8 || i >= (int)out_ptr.size() || s.num_elements() > out_capacity[i]) {
      Status e;
      e.ok_ = false;
      e.msg_ = "harness: output buffer missing or too small";
      return e;
    }
    outputs[i] = Tensor(out_ptr[i], s);
    *t = &outputs[i];
    return Status::OK();
  }
  void SetStatus(const Status& s) { status = s; }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

// REGISTER_OP("X").Input(..).Output(..).Attr(..) -> a throw-away builder object.
struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
};
struct KernelDefBuilderStub {
  KernelDefBuilderStub& Device(const char*) { return *this; }
};
inline KernelDefBuilderStub Name(const char*) { return KernelDefBuilderStub(); }
static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

}  // namespace tensorflow

#define GA_STUB_CAT2(a, b) a##b
#define GA_STUB_CAT(a, b) GA_STUB_CAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderStub GA_STUB_CAT(ga_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub()
#define REGISTER_KERNEL_BUILDER(builder, cls) \
  static int GA_STUB_CAT(ga_stub_kernel_, __COUNTER__) = ((void)(builder), 0)

#define OP_REQUIRES(ctx, cond, status) \
  do {                                 \
    if (!(cond)) {                     \
      (ctx)->SetStatus(status);        \
      return;                          \
    }                                  \
  } while (0)
#define OP_REQUIRES_OK(ctx, expr)          \
  do {                                     \
    ::tensorflow::Status _s = (expr);      \
    if (!_s.ok()) {                        \
      (ctx)->SetStatus(_s);                \
      return;                              \
    }                                      \
  } while (0)
