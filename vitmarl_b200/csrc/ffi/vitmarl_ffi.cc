// XLA FFI (jax.ffi) handlers over the C ABI of include/vitmarl_b200.h -- the reference-side binding for the three call
// sites the hot path replaces:
//   vitmarl_lob_step     job.scan_through_entire_array_save_bidask under vmap   JaxOrderBookArrays.py:720-752 (marl_env.py:377-384)
//   vitmarl_env_step     the same + _ffill_best_prices + mid + get_vision_L2_state / normalize_vision_obs + raster
//                        marl_env.py:392-393,466-467,685-711; JaxOrderBookArrays.py:1108-1140; vision_env.py:2804-2854
//   vitmarl_vit_fwd/bwd  module.apply({'params': p}, x) and its VJP            ippo_rnn_JAXMARL.py:317,423-475
// Thin by design: every handler unpacks buffers / attributes, calls ONE C-ABI entry point on the stream XLA hands in, and
// maps the VITMARL_E* code to ffi::Error.  No state, no allocation (scratch comes from XLA as a result buffer), no sync.
//
// Built only when the XLA FFI headers are present (`python -c "import jax.ffi; print(jax.ffi.include_dir())"`); this image
// ships no JAX (no wheel, no network), so here the translation unit compiles to nothing and vitmarl_b200/_build.py skips it.
// vitmarl_b200/jax_ops.py holds the matching jax.ffi.register_ffi_target / ffi_call / custom_vjp wrappers.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define VITMARL_HAVE_XLA_FFI 1
#endif
#endif

#ifdef VITMARL_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../../include/vitmarl_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int rc, const char* what) {
  if (rc == VITMARL_OK) return ffi::Error::Success();
  const std::string msg = std::string(what) + ": " + (rc == VITMARL_ECUDA ? vitmarl_last_error() : "invalid argument / unsupported mode");
  if (rc == VITMARL_EINVAL || rc == VITMARL_EUNSUPPORTED) return ffi::Error::InvalidArgument(msg);
  return ffi::Error::Internal(msg);
}

// Batch dimension: registered with vmap_method="broadcast_all", so the leading axis of every operand is E.
ffi::Error LobStepImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> msgs, ffi::Buffer<ffi::S32> asks, ffi::Buffer<ffi::S32> bids,
                       ffi::Buffer<ffi::S32> trades, ffi::ResultBuffer<ffi::S32> asks_out, ffi::ResultBuffer<ffi::S32> bids_out,
                       ffi::ResultBuffer<ffi::S32> trades_out, ffi::ResultBuffer<ffi::S32> best_asks,
                       ffi::ResultBuffer<ffi::S32> best_bids, int32_t n_keep, int32_t cancel_mode, int32_t init_id) {
  const auto d = asks.dimensions();                       // [E, N, 6]
  if (d.size() != 3 || msgs.dimensions().size() != 3) return ffi::Error::InvalidArgument("vitmarl_lob_step: expected batched [E,N,6] / [E,M,8]");
  const int E = (int)d[0], N = (int)d[1], M = (int)msgs.dimensions()[1], T = (int)trades.dimensions()[1];
  return to_error(vitmarl_lob_step(stream, E, N, T, M, n_keep, asks.typed_data(), bids.typed_data(), trades.typed_data(), msgs.typed_data(),
                                   asks_out->typed_data(), bids_out->typed_data(), trades_out->typed_data(), best_asks->typed_data(),
                                   best_bids->typed_data(), cancel_mode, init_id),
                  "vitmarl_lob_step");
}

ffi::Error EnvStepImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> msgs, ffi::Buffer<ffi::S32> asks, ffi::Buffer<ffi::S32> bids,
                       ffi::Buffer<ffi::S32> prev_best_asks, ffi::Buffer<ffi::S32> prev_best_bids,
                       ffi::ResultBuffer<ffi::S32> asks_out, ffi::ResultBuffer<ffi::S32> bids_out, ffi::ResultBuffer<ffi::S32> trades_out,
                       ffi::ResultBuffer<ffi::S32> best_asks, ffi::ResultBuffer<ffi::S32> best_bids, ffi::ResultBuffer<ffi::F32> mid_price,
                       ffi::ResultBuffer<ffi::F32> vision_obs, ffi::ResultBuffer<ffi::BF16> image, ffi::ResultBuffer<ffi::S32> trade_stats,
                       int32_t n_levels, int32_t tick_size, int32_t img_h, int32_t img_w, int32_t patch, int32_t cancel_mode, int32_t init_id,
                       ffi::Span<const int32_t> stat_agent_ids) {
  const auto d = asks.dimensions();
  if (d.size() != 3 || msgs.dimensions().size() != 3 || prev_best_asks.dimensions().size() != 3)
    return ffi::Error::InvalidArgument("vitmarl_env_step: expected batched [E,N,6] / [E,M,8] / [E,M,2]");
  VitmarlEnvStepArgs a{};
  a.E = (int)d[0]; a.N = (int)d[1]; a.M = (int)msgs.dimensions()[1]; a.T = (int)trades_out->dimensions()[1];
  a.asks_in = asks.typed_data(); a.bids_in = bids.typed_data(); a.msgs = msgs.typed_data();
  const int Mprev = (int)prev_best_asks.dimensions()[1];   // state.world_state.best_asks[-1, 0] read by stride: no gather op in the graph
  a.last_ask_price = prev_best_asks.typed_data() + (Mprev - 1) * 2;
  a.last_bid_price = prev_best_bids.typed_data() + (Mprev - 1) * 2;
  a.last_price_stride = 2 * Mprev;
  a.asks_out = asks_out->typed_data(); a.bids_out = bids_out->typed_data(); a.trades_out = trades_out->typed_data();
  a.best_asks = best_asks->typed_data(); a.best_bids = best_bids->typed_data(); a.mid_price = mid_price->typed_data();
  a.n_levels = n_levels; a.tick_size = tick_size; a.norm = vision_obs->typed_data();
  a.image = image->untyped_data(); a.H = img_h; a.W = img_w;
  a.img_dtype = patch > 0 ? VITMARL_IMG_BF16_PATCHES_OF(patch) : VITMARL_IMG_BF16;
  a.cancel_mode = cancel_mode; a.init_id = init_id;
  a.n_stat_agents = (int)stat_agent_ids.size();
  if (a.n_stat_agents > 4) return ffi::Error::InvalidArgument("vitmarl_env_step: at most 4 stat_agent_ids");
  for (int i = 0; i < a.n_stat_agents; ++i) a.stat_agent_ids[i] = stat_agent_ids[i];
  a.trade_stats = a.n_stat_agents ? trade_stats->typed_data() : nullptr;
  return to_error(vitmarl_env_step2(stream, &a), "vitmarl_env_step");
}

// params arrive as ONE flat operand per dtype is not how flax lays them out, so the Python wrapper passes the packed table
// (5 + 12 L buffers) as variadic operands; XLA's ScratchAllocator-style workspace is a result buffer sized by
// vitmarl_vit_workspace_bytes (kept as a residual for the backward handler).
ffi::Error VitFwdImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::ResultBuffer<ffi::F32> y, ffi::ResultBuffer<ffi::U8> workspace,
                      int32_t batch, int32_t img_h, int32_t img_w, int32_t channels, int32_t patch, int32_t dim, int32_t depth,
                      int32_t heads, int32_t mlp_dim, float ln_eps, int32_t mode) {
  VitmarlVitShape s{batch, img_h, img_w, channels, patch, dim, depth, heads, mlp_dim, ln_eps};
  const int np = vitmarl_vit_num_params(&s);
  if (np < 0 || (int)args.size() != np + 1) return ffi::Error::InvalidArgument("vitmarl_vit_fwd: expected x followed by the packed parameter table");
  const void* params[5 + 12 * 64];
  if (np > (int)(sizeof(params) / sizeof(params[0]))) return ffi::Error::InvalidArgument("vitmarl_vit_fwd: depth too large");
  auto x = args.get<ffi::AnyBuffer>(0);
  if (!x.has_value()) return ffi::Error::InvalidArgument("vitmarl_vit_fwd: x");
  for (int i = 0; i < np; ++i) {
    auto b = args.get<ffi::AnyBuffer>(i + 1);
    if (!b.has_value()) return ffi::Error::InvalidArgument("vitmarl_vit_fwd: parameter buffer");
    params[i] = b->untyped_data();
  }
  return to_error(vitmarl_vit_fwd(stream, &s, params, x->untyped_data(), y->typed_data(), workspace->typed_data(),
                                  workspace->element_count(), mode),
                  "vitmarl_vit_fwd");
}

ffi::Error VitBwdImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t batch, int32_t img_h, int32_t img_w,
                      int32_t channels, int32_t patch, int32_t dim, int32_t depth, int32_t heads, int32_t mlp_dim, float ln_eps) {
  VitmarlVitShape s{batch, img_h, img_w, channels, patch, dim, depth, heads, mlp_dim, ln_eps};
  const int np = vitmarl_vit_num_params(&s);
  // operands: dy, workspace (the forward's residual; donated), packed params ... ; results: packed gradient table (fp32)
  if (np < 0 || (int)args.size() != np + 2 || (int)rets.size() != np) return ffi::Error::InvalidArgument("vitmarl_vit_bwd: operand / result count");
  const void* params[5 + 12 * 64];
  void* grads[5 + 12 * 64];
  auto dy = args.get<ffi::Buffer<ffi::F32>>(0);
  auto ws = args.get<ffi::Buffer<ffi::U8>>(1);
  if (!dy.has_value() || !ws.has_value()) return ffi::Error::InvalidArgument("vitmarl_vit_bwd: dy / workspace");
  for (int i = 0; i < np; ++i) {
    auto b = args.get<ffi::AnyBuffer>(i + 2);
    auto g = rets.get<ffi::AnyBuffer>(i);
    if (!b.has_value() || !g.has_value()) return ffi::Error::InvalidArgument("vitmarl_vit_bwd: parameter / gradient buffer");
    params[i] = b->untyped_data();
    grads[i] = (*g)->untyped_data();
  }
  return to_error(vitmarl_vit_bwd(stream, &s, params, ws->typed_data(), ws->element_count(), dy->typed_data(), grads, nullptr), "vitmarl_vit_bwd");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitmarlLobStep, LobStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<int32_t>("n_keep").Attr<int32_t>("cancel_mode").Attr<int32_t>("init_id"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitmarlEnvStep, EnvStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::BF16>>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<int32_t>("n_levels").Attr<int32_t>("tick_size").Attr<int32_t>("img_h").Attr<int32_t>("img_w")
                                  .Attr<int32_t>("patch").Attr<int32_t>("cancel_mode").Attr<int32_t>("init_id")
                                  .Attr<ffi::Span<const int32_t>>("stat_agent_ids"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitmarlVitFwd, VitFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("batch").Attr<int32_t>("img_h").Attr<int32_t>("img_w").Attr<int32_t>("channels")
                                  .Attr<int32_t>("patch").Attr<int32_t>("dim").Attr<int32_t>("depth").Attr<int32_t>("heads")
                                  .Attr<int32_t>("mlp_dim").Attr<float>("ln_eps").Attr<int32_t>("mode"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(VitmarlVitBwd, VitBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int32_t>("batch").Attr<int32_t>("img_h").Attr<int32_t>("img_w").Attr<int32_t>("channels")
                                  .Attr<int32_t>("patch").Attr<int32_t>("dim").Attr<int32_t>("depth").Attr<int32_t>("heads")
                                  .Attr<int32_t>("mlp_dim").Attr<float>("ln_eps"));
#endif  // VITMARL_HAVE_XLA_FFI
