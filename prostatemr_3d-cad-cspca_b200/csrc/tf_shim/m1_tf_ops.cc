// m1_tf_ops.cc — TensorFlow 2.5 custom ops over the C-ABI of libm1b200.so (include/m1b200.h).
//
// SOURCE ONLY in this repository: the image has no TensorFlow headers, so this file is not compiled by
// csrc/build.sh and not exercised by the tests (CMakeLists.txt next to it builds it where TF 2.5 is installed).
// It shows how the reference's graph (tf2.5/scripts/model/unets/network_blocks.py) reaches the sm_100a kernels
// without leaving TensorFlow: every op receives DEVICE tensors (TF has already placed them on the GPU), passes
// their raw pointers plus the op's CUDA stream to the library and never synchronises - the same contract the
// ctypes host of this repository uses through DLPack.
//
//   M1InstanceNormLRelu      tfa.layers.InstanceNormalization + LeakyReLU      R:network_blocks.py:38-44,55-58
//   M1InstanceNormLReluGrad  its gradient (registered in Python with tf.RegisterGradient)
//   M1SeGate                 norm3/norm4 + squeeze-excite + gate*residual + LeakyReLU + dropout
//                                                                              R:network_blocks.py:59-78,137-143
//   M1AdamAmsgrad            Keras Adam(amsgrad=True) update + L2              train_model.py:113-120
// (the convolution, attention, latent and loss entry points follow the same pattern - see the note at the end)
//
// Layout contract: NDHWC, fp32 / bf16 / fp16 activations ("T"), fp32 statistics and parameters.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include <mutex>
#include <vector>

#include "m1b200.h"

namespace m1tf {
using namespace tensorflow;  // NOLINT
using GPUDevice = Eigen::GpuDevice;

// one library context per GPU ordinal, created on first use (m1_ctx: TMA encode entry point, scratch)
static m1_ctx* ContextFor(OpKernelContext* c) {
  static std::mutex mu;
  static std::vector<m1_ctx*> ctxs(64, nullptr);
  const int dev = c->device()->tensorflow_gpu_device_info()->gpu_id;
  std::lock_guard<std::mutex> lock(mu);
  if (ctxs[dev] == nullptr && m1_ctx_create(dev, &ctxs[dev]) != 0) return nullptr;
  return ctxs[dev];
}
static void* StreamOf(OpKernelContext* c) { return static_cast<void*>(c->eigen_device<GPUDevice>().stream()); }
template <typename T> struct DType;
template <> struct DType<float> { static constexpr int v = M1_F32; };
template <> struct DType<bfloat16> { static constexpr int v = M1_BF16; };
template <> struct DType<Eigen::half> { static constexpr int v = M1_F16; };
static const void* Ptr(const Tensor& t) { return t.tensor_data().data(); }
static void* Ptr(Tensor* t) { return const_cast<char*>(t->tensor_data().data()); }
#define M1_OK(c, expr) OP_REQUIRES((c), (expr) == 0, errors::Internal("libm1b200: ", m1_last_error()))

static void Nvc(const Tensor& x, int* n, int64_t* v, int* ch) {      // (B, D, H, W, C) -> batch, voxels, channels
  *n = static_cast<int>(x.dim_size(0));
  *ch = static_cast<int>(x.dim_size(x.dims() - 1));
  *v = x.NumElements() / (static_cast<int64_t>(*n) * *ch);
}

// ---- InstanceNorm (eps 1e-3, biased variance, affine) + LeakyReLU(slope) ------------------------------------
REGISTER_OP("M1InstanceNormLRelu")
    .Input("x: T").Input("gamma: float").Input("beta: float")
    .Attr("slope: float = 0.1").Attr("epsilon: float = 0.001").Attr("T: {float, bfloat16, half}")
    .Output("y: T").Output("stats: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->UnknownShape());
      return Status::OK();
    });
template <typename T>
class InstanceNormLReluOp : public OpKernel {
 public:
  explicit InstanceNormLReluOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("slope", &slope_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor& x = c->input(0);
    int n, ch; int64_t v;
    Nvc(x, &n, &v, &ch);
    Tensor *y = nullptr, *stats = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &y));
    OP_REQUIRES_OK(c, c->allocate_output(1, TensorShape({n, ch, 2}), &stats));
    m1_ctx* ctx = ContextFor(c);
    OP_REQUIRES(c, ctx != nullptr, errors::Internal("libm1b200: ", m1_last_error()));
    M1_OK(c, m1_inorm_stats(ctx, Ptr(x), DType<T>::v, n, v, ch, eps_, stats->flat<float>().data(), StreamOf(c)));
    M1_OK(c, m1_inorm_act_fwd(ctx, Ptr(x), stats->flat<float>().data(), c->input(1).flat<float>().data(),
                              c->input(2).flat<float>().data(), DType<T>::v, n, v, ch, slope_, Ptr(y), nullptr,
                              StreamOf(c)));
  }
 private:
  float slope_, eps_;
};

REGISTER_OP("M1InstanceNormLReluGrad")
    .Input("dy: T").Input("x: T").Input("stats: float").Input("gamma: float").Input("beta: float")
    .Attr("slope: float = 0.1").Attr("T: {float, bfloat16}")
    .Output("dx: T").Output("dgamma: float").Output("dbeta: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(1));
      c->set_output(1, c->input(3));
      c->set_output(2, c->input(4));
      return Status::OK();
    });
template <typename T>
class InstanceNormLReluGradOp : public OpKernel {
 public:
  explicit InstanceNormLReluGradOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("slope", &slope_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor& x = c->input(1);
    int n, ch; int64_t v;
    Nvc(x, &n, &v, &ch);
    Tensor *dx = nullptr, *dg = nullptr, *db = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(c, c->allocate_output(1, c->input(3).shape(), &dg));
    OP_REQUIRES_OK(c, c->allocate_output(2, c->input(4).shape(), &db));
    auto stream = c->eigen_device<GPUDevice>().stream();
    cudaMemsetAsync(dg->flat<float>().data(), 0, sizeof(float) * ch, stream);      // the kernel accumulates
    cudaMemsetAsync(db->flat<float>().data(), 0, sizeof(float) * ch, stream);
    m1_ctx* ctx = ContextFor(c);
    OP_REQUIRES(c, ctx != nullptr, errors::Internal("libm1b200: ", m1_last_error()));
    M1_OK(c, m1_inorm_act_bwd(ctx, Ptr(c->input(0)), Ptr(x), c->input(2).flat<float>().data(),
                              c->input(3).flat<float>().data(), c->input(4).flat<float>().data(), DType<T>::v, n, v, ch,
                              slope_, Ptr(dx), /*accumulate=*/0, dg->flat<float>().data(), db->flat<float>().data(),
                              StreamOf(c)));
  }
 private:
  float slope_;
};

// ---- SE tail: out = dropout(lrelu(norm3(raw3) * sigmoid(conv7(lrelu(conv6(GAP(norm3(raw3)))))) * norm4(raw4))) ----
REGISTER_OP("M1SeGate")
    .Input("raw3: T").Input("raw4: T")
    .Input("gamma3: float").Input("beta3: float").Input("gamma4: float").Input("beta4: float")
    .Input("w6: float").Input("b6: float").Input("w7: float").Input("b7: float")
    .Attr("rate: float = 0.0").Attr("seed: int = 42").Attr("stream_id: int = 0").Attr("T: {float, bfloat16, half}")
    .Output("out: T").Output("stats3: float").Output("stats4: float").Output("gate: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      for (int i = 1; i < 4; ++i) c->set_output(i, c->UnknownShape());
      return Status::OK();
    });
template <typename T>
class SeGateOp : public OpKernel {
 public:
  explicit SeGateOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("rate", &rate_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
    OP_REQUIRES_OK(c, c->GetAttr("stream_id", &stream_id_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor& raw3 = c->input(0);
    int n, ch; int64_t v;
    Nvc(raw3, &n, &v, &ch);
    const int cr = static_cast<int>(c->input(7).NumElements());
    Tensor *out, *st3, *st4, *gate, pool, hidden;
    OP_REQUIRES_OK(c, c->allocate_output(0, raw3.shape(), &out));
    OP_REQUIRES_OK(c, c->allocate_output(1, TensorShape({n, ch, 2}), &st3));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({n, ch, 2}), &st4));
    OP_REQUIRES_OK(c, c->allocate_output(3, TensorShape({n, ch}), &gate));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_FLOAT, TensorShape({n, ch}), &pool));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_FLOAT, TensorShape({n, cr}), &hidden));
    m1_ctx* ctx = ContextFor(c);
    OP_REQUIRES(c, ctx != nullptr, errors::Internal("libm1b200: ", m1_last_error()));
    void* s = StreamOf(c);
    const float *g3 = c->input(2).flat<float>().data(), *b3 = c->input(3).flat<float>().data(),
                *g4 = c->input(4).flat<float>().data(), *b4 = c->input(5).flat<float>().data();
    M1_OK(c, m1_inorm_stats(ctx, Ptr(raw3), DType<T>::v, n, v, ch, 1e-3f, st3->flat<float>().data(), s));
    M1_OK(c, m1_inorm_stats(ctx, Ptr(c->input(1)), DType<T>::v, n, v, ch, 1e-3f, st4->flat<float>().data(), s));
    M1_OK(c, m1_se_excite_fwd(ctx, pool.flat<float>().data(), c->input(6).flat<float>().data(),
                              c->input(7).flat<float>().data(), c->input(8).flat<float>().data(),
                              c->input(9).flat<float>().data(), n, ch, cr, hidden.flat<float>().data(),
                              gate->flat<float>().data(), st3->flat<float>().data(), g3, b3, s));
    m1_dropout drop = {};
    drop.rate = rate_;
    drop.seed = static_cast<uint64_t>(seed_);
    drop.stream_id = static_cast<uint64_t>(stream_id_);
    M1_OK(c, m1_se_gate_fwd(ctx, Ptr(raw3), Ptr(c->input(1)), st3->flat<float>().data(), st4->flat<float>().data(), g3, b3,
                            g4, b4, gate->flat<float>().data(), &drop, DType<T>::v, n, v, ch, Ptr(out), nullptr, s));
  }
 private:
  float rate_;
  int64 seed_, stream_id_;
};

// ---- Keras Adam(amsgrad=True) + L2 regulariser, one fused pass over w, g, m, v, v-hat -----------------------------
REGISTER_OP("M1AdamAmsgrad")
    .Input("w: Ref(float)").Input("g: float").Input("m: Ref(float)").Input("v: Ref(float)").Input("vhat: Ref(float)")
    .Input("lr_t: float")
    .Attr("beta_1: float = 0.9").Attr("beta_2: float = 0.999").Attr("epsilon: float = 1e-7").Attr("l2: float = 0.0")
    .Output("l2_loss: float")
    .SetShapeFn(shape_inference::ScalarShape);
class AdamAmsgradOp : public OpKernel {
 public:
  explicit AdamAmsgradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("beta_1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("beta_2", &b2_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("l2", &l2_));
  }
  void Compute(OpKernelContext* c) override {
    Tensor w = c->mutable_input(0, true), m = c->mutable_input(2, true), v = c->mutable_input(3, true),
           vh = c->mutable_input(4, true);
    Tensor* l2 = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &l2));
    cudaMemsetAsync(l2->flat<float>().data(), 0, sizeof(float), c->eigen_device<GPUDevice>().stream());
    m1_ctx* ctx = ContextFor(c);
    OP_REQUIRES(c, ctx != nullptr, errors::Internal("libm1b200: ", m1_last_error()));
    // the step size lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) arrives as a DEVICE scalar: graph-replay friendly
    M1_OK(c, m1_adam_amsgrad_dev(ctx, w.flat<float>().data(), c->input(1).flat<float>().data(), m.flat<float>().data(),
                                 v.flat<float>().data(), vh.flat<float>().data(), w.NumElements(),
                                 c->input(5).flat<float>().data(), b1_, b2_, eps_, l2_, 1.0f, l2->flat<float>().data(),
                                 /*amsgrad=*/1, StreamOf(c)));
  }
 private:
  float b1_, b2_, eps_, l2_;
};

#define REGISTER_M1_T(T)                                                                                      \
  REGISTER_KERNEL_BUILDER(Name("M1InstanceNormLRelu").Device(DEVICE_GPU).TypeConstraint<T>("T"),             \
                          InstanceNormLReluOp<T>);                                                            \
  REGISTER_KERNEL_BUILDER(Name("M1SeGate").Device(DEVICE_GPU).TypeConstraint<T>("T"), SeGateOp<T>);
REGISTER_M1_T(float)
REGISTER_M1_T(bfloat16)
REGISTER_M1_T(Eigen::half)
REGISTER_KERNEL_BUILDER(Name("M1InstanceNormLReluGrad").Device(DEVICE_GPU).TypeConstraint<float>("T"),
                        InstanceNormLReluGradOp<float>);
REGISTER_KERNEL_BUILDER(Name("M1InstanceNormLReluGrad").Device(DEVICE_GPU).TypeConstraint<bfloat16>("T"),
                        InstanceNormLReluGradOp<bfloat16>);
REGISTER_KERNEL_BUILDER(Name("M1AdamAmsgrad").Device(DEVICE_GPU).HostMemory("l2_loss"), AdamAmsgradOp);

// The convolution (m1_conv3d + m1_conv3d_pack_weights), attention (m1_attn_fwd / _bwd), latent (m1_latent_fwd,
// m1_kl_fwd) and loss (m1_logits_softmax_focal) ops follow the same pattern: attrs carry the m1_conv_desc fields
// (kernel, stride, transposed flag, fused output split), list(T) inputs the virtual concatenation, and the packed
// 16-bit weight operand is a persistent tensor re-derived after every optimizer step (Engine.refresh_packs).
}  // namespace m1tf
