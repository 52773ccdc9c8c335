// nqmc_api.cu -- the C ABI of include/nqmc_b200.h: context, parameter map, entry points.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nqmc_internal.cuh"


static std::string g_create_err;

namespace {

struct LeafBuilder {
  std::vector<LeafInfo>& leaves;
  int64_t off = 0;
  int add(const std::string& path, int64_t rows, int64_t cols) {
    leaves.push_back({path, rows, cols, off});
    int o = (int)off;
    off += rows * cols;
    return o;
  }
  // flax Dense: 'bias' (out,) sorts before 'kernel' (in,out)
  void dense(const std::string& prefix, int in, int out, int& b, int& k) {
    b = add(prefix + "/bias", 1, out);
    k = add(prefix + "/kernel", in, out);
  }
  MlpOff mlp3(const std::string& prefix, int F, int H) {
    MlpOff o;
    dense(prefix + "/Dense_0", F, H, o.b1, o.W1);
    dense(prefix + "/Dense_1", H, H, o.b2, o.W2);
    dense(prefix + "/Dense_2", H, 1, o.b3, o.W3);
    return o;
  }
};

// Builds the sorted-key (jax.tree.leaves) parameter map.  Mirrors oracle/layout.py::leaf_table,
// which tests/test_abi_layout.py checks it against.
void build_param_map(nqmc_ctx* ctx) {
  DevSys& s = ctx->sys;
  LeafBuilder lb{ctx->leaves};
  if (s.kind == NQMC_KIND_SIMPLE) {
    // Dense_0 ... Dense_n in sorted-key order ("Dense_10" would sort before "Dense_2"; n <= 6 here)
    for (int l = 0; l <= s.simple.n_hidden; ++l) {
      const int out = l == s.simple.n_hidden ? 1 : s.simple.dim[l + 1];
      lb.dense("params/Dense_" + std::to_string(l), s.simple.dim[l], out, s.simple.offb[l], s.simple.offW[l]);
    }
    s.P = (int)lb.off;
    return;
  }
  if (ctx->want_bf) s.bf = lb.mlp3("params/backflow", 2, NQMC_BF_H);
  if (s.use_cusp) {
    s.ee_cusp_b = lb.add("params/ee_cusp/b_ee", 1, 1);
    std::vector<std::pair<std::string, int>> names;
    for (int d = 0; d < 2 * s.A; ++d) names.push_back({"Dense_" + std::to_string(d), d});
    std::sort(names.begin(), names.end());
    for (auto& nd : names) {
      const int d = nd.second, a = d / 2;
      if (d % 2 == 0) lb.dense("params/en_cusp/" + nd.first, 3, NQMC_CUSP_H, s.en_cusp_b0[a], s.en_cusp_W0[a]);
      else lb.dense("params/en_cusp/" + nd.first, NQMC_CUSP_H, 1, s.en_cusp_b1[a], s.en_cusp_W1[a]);
    }
  }
  s.jas_a_ee = lb.add("params/jastrow/a_ee", 1, 1);
  s.jas_a_en = lb.add("params/jastrow/a_en", 1, 1);
  s.jas_b_ee = lb.add("params/jastrow/b_ee", 1, 1);
  s.jas_b_en = lb.add("params/jastrow/b_en", 1, 1);
  for (int k = 0; k < s.n_dn; ++k)
    s.mlp[s.n_up + k] = lb.mlp3("params/orbitals_down/orbital_" + std::to_string(k), s.F, s.H);
  for (int k = 0; k < s.n_up; ++k) s.mlp[k] = lb.mlp3("params/orbitals_up/orbital_" + std::to_string(k), s.F, s.H);
  s.P = (int)lb.off;
}

inline cudaStream_t S(nqmc_stream st) { return reinterpret_cast<cudaStream_t>(st); }

int check_w(nqmc_ctx* ctx, int64_t W) {
  if (!ctx) return NQMC_ERR_ARG;
  if (W <= 0) {
    ctx->err = "W must be positive";
    return NQMC_ERR_ARG;
  }
  if (W > ctx->cap_w) {
    ctx->err = "W exceeds the reserved capacity; call nqmc_reserve first";
    return NQMC_ERR_STATE;
  }
  return NQMC_OK;
}

// Scratch allocation.  With NQMC_GUARD=1 in the environment at nqmc_create every scratch buffer gets a 4 KB guard band on
// either side, filled with a pattern that nqmc_debug_check_guards() verifies: the out-of-bounds check of this library
// (compute-sanitizer is not available on the GPU pool this was developed on).
constexpr size_t GUARD_BYTES = 4096;
constexpr int GUARD_BYTE = 0xA5;
cudaError_t scratch_alloc(nqmc_ctx* c, void** out, size_t bytes, const char* name) {
  if (!c->guard) return cudaMalloc(out, bytes);
  const size_t padded = (bytes + 255) / 256 * 256;
  char* base = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), padded + 2 * GUARD_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaMemset(base, GUARD_BYTE, padded + 2 * GUARD_BYTES);
  if (e != cudaSuccess) return e;
  *out = base + GUARD_BYTES;
  c->guards.push_back({base, bytes, padded, name});
  return cudaSuccess;
}
void scratch_free(nqmc_ctx* c, void* p) {
  if (!p) return;
  if (c->guard) {
    for (auto& g : c->guards)
      if (g.base + GUARD_BYTES == p) { cudaFree(g.base); g.base = nullptr; return; }
  }
  cudaFree(p);
}

void free_scratch(nqmc_ctx* c) {
  void* ptrs[] = {c->bfw, c->gbf, c->x_cur, c->x_prop, c->vvec, c->r_prop, c->phi, c->inv, c->logabs, c->sign, c->e_scratch, c->wgt, c->r_soa, c->r_soa2, c->nrm_soa,
                  c->aos_stage, c->uni_stage, c->logp_stage, c->nacc_stage, c->hdr_stage, c->gsum_stage, c->part_walk,
                  c->part_mlp, c->hdr_tmp, c->jas_sums, c->bw_h, c->bw_d, c->bw_e, c->bw_f, c->bw_h2, c->l1cache};
  for (void* p : ptrs) scratch_free(c, p);
  c->guards.clear();
  c->x_cur = c->x_prop = c->vvec = c->gbf = c->bfw = nullptr;
  c->r_prop = c->phi = c->inv = c->logabs = c->sign = c->e_scratch = c->wgt = c->r_soa = c->r_soa2 = c->nrm_soa = nullptr;
  c->up_pending = false;
  c->aos_stage = c->uni_stage = c->logp_stage = c->nacc_stage = nullptr;
  c->hdr_stage = c->gsum_stage = c->part_walk = c->hdr_tmp = c->jas_sums = nullptr;
  c->jas_valid_for = nullptr;
  c->part_mlp = nullptr;
  c->bw_h = c->bw_d = c->bw_e = c->bw_f = c->bw_h2 = nullptr;
  c->act_valid_for = nullptr;
  c->l1cache = nullptr;
  c->bw_rows = 0;
  c->cap_w = 0;
}

}  // namespace

extern "C" {

int nqmc_abi_version(void) { return NQMC_ABI_VERSION; }

const char* nqmc_last_error(const nqmc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int nqmc_create(const nqmc_system* sys, const nqmc_net* net, int device, nqmc_ctx** out) {
  if (!sys || !net || !out) {
    g_create_err = "null argument";
    return NQMC_ERR_ARG;
  }
  *out = nullptr;
  if (sys->n_atoms < 1 || sys->n_atoms > NQMC_MAX_ATOMS || sys->n_elec < 1 || sys->n_elec > NQMC_MAX_ELEC ||
      sys->n_up < 0 || sys->n_up > sys->n_elec || !sys->R || !sys->Z) {
    g_create_err = "system out of range (1..16 atoms, 1..16 electrons, 0 <= n_up <= n_elec)";
    return NQMC_ERR_ARG;
  }
  const int n_dn = sys->n_elec - sys->n_up;
  if (sys->n_up > NQMC_MAX_NS || n_dn > NQMC_MAX_NS) {
    g_create_err = "more than 8 electrons per spin are not supported";
    return NQMC_ERR_UNSUPPORTED;
  }
  if (net->kind != NQMC_KIND_SIMPLE && net->kind != NQMC_KIND_ANTISYMMETRIC) {
    g_create_err = "unknown wavefunction kind";
    return NQMC_ERR_ARG;
  }
  if (net->kind == NQMC_KIND_ANTISYMMETRIC && net->hidden0 != net->hidden1) {
    g_create_err = "orbital hidden_dims must be two equal widths";
    return NQMC_ERR_UNSUPPORTED;
  }
  if (net->kind == NQMC_KIND_ANTISYMMETRIC) {
    if (net->hidden0 != 32 && net->hidden0 != 64 && net->hidden0 != 128) {
      g_create_err = "orbital hidden width must be 32, 64 or 128";
      return NQMC_ERR_UNSUPPORTED;
    }
    if (3 + 2 * sys->n_atoms > net->hidden0) {
      g_create_err = "3 + 2*n_atoms orbital features must not exceed the hidden width";
      return NQMC_ERR_UNSUPPORTED;
    }
  } else {
    if (net->n_hidden < 0 || net->n_hidden > NQMC_SIMPLE_MAX_HIDDEN) {
      g_create_err = "SimpleWavefunction: at most 6 hidden layers";
      return NQMC_ERR_UNSUPPORTED;
    }
    if (net->activation != NQMC_ACT_TANH && net->activation != NQMC_ACT_SILU) {
      g_create_err = "Unknown activation";
      return NQMC_ERR_ARG;
    }
    const int nh = net->n_hidden ? net->n_hidden : 2;
    for (int l = 0; l < nh; ++l) {
      const int h = net->n_hidden ? net->hidden[l] : (l == 0 ? net->hidden0 : net->hidden1);
      if (h < 1 || h > 128) {
        g_create_err = "SimpleWavefunction hidden widths must be in 1..128";
        return NQMC_ERR_UNSUPPORTED;
      }
    }
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return NQMC_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_err = "device index out of range";
    return NQMC_ERR_ARG;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return NQMC_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_err = "device is not sm_100 (this library is built for B200 only)";
    return NQMC_ERR_CUDA;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return NQMC_ERR_CUDA;
  }
  nqmc_ctx* ctx = new nqmc_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  DevSys& s = ctx->sys;
  memset(&s, 0, sizeof(s));
  s.A = sys->n_atoms;
  s.N = sys->n_elec;
  s.n_up = sys->n_up;
  s.n_dn = n_dn;
  s.kind = net->kind;
  s.H = net->hidden0;
  s.use_cusp = net->kind == NQMC_KIND_ANTISYMMETRIC ? (net->use_cusp != 0) : 0;
  s.use_bf = net->kind == NQMC_KIND_ANTISYMMETRIC ? (net->use_backflow != 0 && sys->n_elec >= 2) : 0;
  s.F = net->kind == NQMC_KIND_ANTISYMMETRIC ? 3 + 2 * s.A : 3 * s.N;
  s.n_mlp = s.n_up + s.n_dn;
  s.n_ent = s.n_up * s.n_up + s.n_dn * s.n_dn;
  s.v_nn = (float)sys->v_nn;
  s.env_alpha = (float)net->envelope_decay;
  s.cusp_same = (float)net->cusp_same;
  s.cusp_diff = (float)net->cusp_diff;
  for (int a = 0; a < s.A; ++a) {
    for (int d = 0; d < 3; ++d) s.R[3 * a + d] = (float)sys->R[3 * a + d];
    s.Z[a] = (float)sys->Z[a];
  }
  if (net->kind == NQMC_KIND_SIMPLE) {
    s.simple.n_hidden = net->n_hidden ? net->n_hidden : 2;
    s.simple.act = net->activation;
    s.simple.dim[0] = 3 * s.N;
    for (int l = 0; l < s.simple.n_hidden; ++l)
      s.simple.dim[l + 1] = net->n_hidden ? net->hidden[l] : (l == 0 ? net->hidden0 : net->hidden1);
  }
  if (const char* e = getenv("NQMC_TC")) ctx->use_tc = atoi(e) != 0;
  ctx->want_bf = net->kind == NQMC_KIND_ANTISYMMETRIC && net->use_backflow != 0;
  build_param_map(ctx);
  ctx->p_mlp = s.kind == NQMC_KIND_ANTISYMMETRIC ? (s.F * s.H + s.H + s.H * s.H + s.H + s.H + 1) : s.P;
  e = cudaMalloc(&ctx->theta, sizeof(float) * (size_t)s.P);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaMalloc(theta): ") + cudaGetErrorString(e);
    delete ctx;
    return NQMC_ERR_CUDA;
  }
  cudaMemset(ctx->theta, 0, sizeof(float) * (size_t)s.P);
  *out = ctx;
  return NQMC_OK;
}

int nqmc_destroy(nqmc_ctx* ctx) {
  if (!ctx) return NQMC_ERR_ARG;
  cudaSetDevice(ctx->device);
  free_scratch(ctx);
  if (ctx->theta) cudaFree(ctx->theta);
  nq_comm_free(ctx);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->ev_phase_a) cudaEventDestroy(ctx->ev_phase_a);
  if (ctx->ev_mh_done) cudaEventDestroy(ctx->ev_mh_done);
  if (ctx->ev_side_done) cudaEventDestroy(ctx->ev_side_done);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  if (ctx->ev_upload) cudaEventDestroy(ctx->ev_upload);
  if (ctx->ev_all_done) cudaEventDestroy(ctx->ev_all_done);
  for (int k = 0; k < nqmc_ctx::PROF_KINDS; ++k)
    for (int i = 0; i < nqmc_ctx::PROF_RING; ++i)
      for (int e = 0; e < 2; ++e)
        if (ctx->prof_ev[k][i][e]) cudaEventDestroy(ctx->prof_ev[k][i][e]);
  delete ctx;
  return NQMC_OK;
}

int nqmc_reserve(nqmc_ctx* ctx, int64_t W) {
  if (!ctx || W <= 0) return NQMC_ERR_ARG;
  if (const char* e = getenv("NQMC_GUARD")) ctx->guard = atoi(e) != 0;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  if (W <= ctx->cap_w) return NQMC_OK;
  free_scratch(ctx);
  const DevSys& s = ctx->sys;
  const size_t nn = 3 * (size_t)s.N, w = (size_t)W, f = sizeof(float);
  const size_t n_ent = s.n_ent > 0 ? s.n_ent : 1;
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->r_prop), nn * w * f, "r_prop"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->phi), (s.use_bf ? 10 : 5) * n_ent * w * f, "phi"));
  if (s.use_bf) {
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->x_cur), nn * w * f, "x_cur"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->x_prop), nn * w * f, "x_prop"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->vvec), nn * w * f, "vvec"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->gbf), 2 * nn * w * f, "gbf"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->bfw), nq_backflow_ws_floats(s) * w * f, "bfw"));
  }
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->inv), n_ent * w * f, "inv"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->logabs), w * f, "logabs"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->sign), w * f, "sign"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->e_scratch), w * f, "e_scratch"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->wgt), w * f, "wgt"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->r_soa), nn * w * f, "r_soa"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->r_soa2), nn * w * f, "r_soa2"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->nrm_soa), nn * w * f, "nrm_soa"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->aos_stage), nn * w * f, "aos_stage"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->uni_stage), w * f, "uni_stage"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->logp_stage), w * f, "logp_stage"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->nacc_stage), w * f, "nacc_stage"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->hdr_stage), NQMC_HDR_SIZE * sizeof(double), "hdr_stage"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->hdr_tmp), NQMC_HDR_SIZE * sizeof(double), "hdr_tmp"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->jas_sums), 8 * sizeof(double), "jas_sums"));
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->gsum_stage), (size_t)s.P * sizeof(double), "gsum_stage"));
  ctx->n_blk_walk = 65536;
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->part_walk), (size_t)ctx->n_blk_walk * 16 * sizeof(double), "part_walk"));
  ctx->n_cta_bwd = 2 * ctx->sm_count;
  NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->part_mlp), (size_t)ctx->n_cta_bwd * ctx->p_mlp * f, "part_mlp"));
  if (s.kind == NQMC_KIND_ANTISYMMETRIC && s.H == 128)
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->l1cache), (size_t)ctx->sm_count * 16 * 10 * 512 * f, "l1cache"));
  if (s.kind == NQMC_KIND_ANTISYMMETRIC && s.H == 128 && s.F + 1 <= 32) {
    ctx->bw_rows = nq_bwd128_rows(ctx, W);
    const size_t per = (size_t)s.n_mlp * 128 * (size_t)ctx->bw_rows * f;
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->bw_h), per, "bw_h"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->bw_d), per, "bw_d"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->bw_e), per, "bw_e"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->bw_h2), per, "bw_h2"));
    NQMC_CUDA(scratch_alloc(ctx, reinterpret_cast<void**>(&ctx->bw_f), per / 4, "bw_f"));
    NQMC_CUDA(cudaMemset(ctx->bw_f, 0, per / 4));       // feature rows F+1..31 stay zero
  }
  ctx->cap_w = W;
  return NQMC_OK;
}

int64_t nqmc_param_count(const nqmc_ctx* ctx) { return ctx ? ctx->sys.P : -1; }

int64_t nqmc_leaf_info(const nqmc_ctx* ctx, int index, char* path, size_t cap, int64_t* rows, int64_t* cols,
                       int64_t* offset) {
  if (!ctx) return NQMC_ERR_ARG;
  if (index < 0) return (int64_t)ctx->leaves.size();
  if (index >= (int)ctx->leaves.size()) return NQMC_ERR_ARG;
  const LeafInfo& l = ctx->leaves[index];
  if (path && cap > 0) {
    strncpy(path, l.path.c_str(), cap - 1);
    path[cap - 1] = 0;
  }
  if (rows) *rows = l.rows;
  if (cols) *cols = l.cols;
  if (offset) *offset = l.offset;
  return NQMC_OK;
}

int nqmc_set_params(nqmc_ctx* ctx, const float* theta, int64_t P, nqmc_stream st) {
  if (!ctx || !theta) return NQMC_ERR_ARG;
  if (P != ctx->sys.P) {
    ctx->err = "theta has the wrong length";
    return NQMC_ERR_ARG;
  }
  NQMC_CUDA(cudaSetDevice(ctx->device));
  ctx->theta_dirty = true;
  ctx->act_valid_for = nullptr;
  NQMC_CUDA(cudaMemcpyAsync(ctx->theta, theta, sizeof(float) * (size_t)P, cudaMemcpyDeviceToDevice, S(st)));
  return NQMC_OK;
}

int nqmc_set_params_host(nqmc_ctx* ctx, const float* theta, int64_t P, nqmc_stream st) {
  if (!ctx || !theta) return NQMC_ERR_ARG;
  if (P != ctx->sys.P) {
    ctx->err = "theta has the wrong length";
    return NQMC_ERR_ARG;
  }
  NQMC_CUDA(cudaSetDevice(ctx->device));
  ctx->theta_dirty = true;
  ctx->act_valid_for = nullptr;
  NQMC_CUDA(cudaMemcpyAsync(ctx->theta, theta, sizeof(float) * (size_t)P, cudaMemcpyHostToDevice, S(st)));
  NQMC_CUDA(cudaStreamSynchronize(S(st)));
  return NQMC_OK;
}

int nqmc_get_params(nqmc_ctx* ctx, float* theta, int64_t P, nqmc_stream st) {
  if (!ctx || !theta) return NQMC_ERR_ARG;
  if (P != ctx->sys.P) {
    ctx->err = "theta has the wrong length";
    return NQMC_ERR_ARG;
  }
  NQMC_CUDA(cudaSetDevice(ctx->device));
  NQMC_CUDA(cudaMemcpyAsync(theta, ctx->theta, sizeof(float) * (size_t)P, cudaMemcpyDeviceToDevice, S(st)));
  return NQMC_OK;
}

int nqmc_get_params_host(nqmc_ctx* ctx, float* theta, int64_t P, nqmc_stream st) {
  if (!ctx || !theta) return NQMC_ERR_ARG;
  if (P != ctx->sys.P) {
    ctx->err = "theta has the wrong length";
    return NQMC_ERR_ARG;
  }
  NQMC_CUDA(cudaSetDevice(ctx->device));
  NQMC_CUDA(cudaMemcpyAsync(theta, ctx->theta, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost, S(st)));
  NQMC_CUDA(cudaStreamSynchronize(S(st)));
  return NQMC_OK;
}

int nqmc_adam_step(nqmc_ctx* ctx, const double* gsum, const double* hdr, float* m, float* v, int64_t count, float lr,
                   float clip, double* grad_norm, nqmc_stream st) {
  if (!ctx || !gsum || !hdr || !m || !v || count < 1) return NQMC_ERR_ARG;
  if (ctx->cap_w <= 0) {
    ctx->err = "call nqmc_reserve first";
    return NQMC_ERR_STATE;
  }
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_adam_step(ctx, gsum, hdr, m, v, count, lr, clip, grad_norm, S(st));
}

int nqmc_aos_to_soa(nqmc_ctx* ctx, const float* aos, float* soa, int64_t W, nqmc_stream st) {
  if (!ctx || !aos || !soa || W <= 0) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_aos_soa(ctx, aos, soa, W, true, S(st));
}
int nqmc_soa_to_aos(nqmc_ctx* ctx, const float* soa, float* aos, int64_t W, nqmc_stream st) {
  if (!ctx || !aos || !soa || W <= 0) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_aos_soa(ctx, soa, aos, W, false, S(st));
}

int nqmc_log_psi(nqmc_ctx* ctx, const float* r, int64_t W, float* logabs, float* sign, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !logabs) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->sys.kind == NQMC_KIND_SIMPLE) {
    rc = nq_simple_logpsi(ctx, r, W, logabs, S(st));
    if (rc) return rc;
    if (sign) rc = nq_fill(ctx, sign, 1.0f, W, S(st));   // psi = exp(real) > 0
    return rc;
  }
  const float* pos = r;
  if (ctx->sys.use_bf) {
    rc = nq_launch_backflow_x(ctx, r, W, ctx->x_cur, S(st));
    if (rc) return rc;
    pos = ctx->x_cur;
  }
  rc = nq_launch_orb_eval(ctx, pos, W, 1, ctx->phi, S(st));
  if (rc) return rc;
  return nq_launch_assemble(ctx, r, W, nullptr, nullptr, false, logabs, sign, S(st));
}

static int mh_core(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed,
                   uint64_t step, int64_t wid0, float sigma, int64_t W, uint8_t* accept, float* n_acc,
                   cudaStream_t st) {
  int rc = nq_launch_propose(ctx, r, normals, seed, step, wid0, sigma, W, ctx->r_prop, st);
  if (rc) return rc;
  if (ctx->sys.kind == NQMC_KIND_SIMPLE) {
    rc = nq_simple_logpsi(ctx, ctx->r_prop, W, ctx->logabs, st);
    if (rc) return rc;
    return nq_accept_logabs(ctx, r, logp, ctx->r_prop, ctx->logabs, uniforms, seed, step, wid0, W, accept, n_acc, st);
  }
  const float* pos = ctx->r_prop;
  if (ctx->sys.use_bf) {
    rc = nq_launch_backflow_x(ctx, ctx->r_prop, W, ctx->x_prop, st);
    if (rc) return rc;
    pos = ctx->x_prop;
  }
  rc = nq_launch_orb_eval(ctx, pos, W, 1, ctx->phi, st);
  if (rc) return rc;
  return nq_launch_accept(ctx, r, logp, ctx->r_prop, uniforms, seed, step, wid0, W, accept, n_acc, st);
}

int nqmc_mh_step(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed,
                 uint64_t step, int64_t wid0, float sigma, int64_t W, uint8_t* accept, float* n_acc, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !logp) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return mh_core(ctx, r, logp, normals, uniforms, seed, step, wid0, sigma, W, accept, n_acc, S(st));
}

int nqmc_sample(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed_burn,
                uint64_t seed_sample, int64_t wid0, float sigma, int64_t W, int64_t n_burn, int64_t n_samples,
                int64_t n_thin, float* samples, float* n_acc, float* sigma_dev, float target_acceptance,
                int64_t adapt_every, nqmc_stream st_) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !logp || n_burn < 0 || n_samples < 0 || n_thin < 1 || (n_samples > 0 && !samples)) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = S(st_);
  const size_t nn = 3 * (size_t)ctx->sys.N, w = (size_t)W;
  float* cnt = n_acc ? n_acc : ctx->nacc_stage;
  const bool adapt = sigma_dev != nullptr && adapt_every > 0;
  ctx->sigma_dev = sigma_dev;
  struct Reset { nqmc_ctx* c; ~Reset() { c->sigma_dev = nullptr; } } reset{ctx};
  if ((rc = nq_fill(ctx, cnt, 0.f, W, st))) return rc;                        // metropolis.py:177-182
  int64_t t_all = 0;
  for (int64_t t = 0; t < n_burn; ++t, ++t_all) {                             // metropolis.py:185-191
    rc = mh_core(ctx, r, logp, normals ? normals + t_all * nn * w : nullptr, uniforms ? uniforms + t_all * w : nullptr,
                 seed_burn, (uint64_t)t, wid0, sigma, W, nullptr, cnt, st);
    if (rc) return rc;
    if (adapt && (t + 1) % adapt_every == 0) {
      const float gain = 1.0f / sqrtf((float)((t + 1) / adapt_every));
      if ((rc = nq_adapt_sigma(ctx, cnt, W, (int)adapt_every, target_acceptance, gain, sigma_dev, st))) return rc;
    }
  }
  if ((rc = nq_fill(ctx, cnt, 0.f, W, st))) return rc;                        // metropolis.py:194-199
  for (int64_t t = 0; t < n_samples * n_thin; ++t, ++t_all) {                 // metropolis.py:202-210
    rc = mh_core(ctx, r, logp, normals ? normals + t_all * nn * w : nullptr, uniforms ? uniforms + t_all * w : nullptr,
                 seed_sample, (uint64_t)t, wid0, sigma, W, nullptr, cnt, st);
    if (rc) return rc;
    if (t % n_thin == 0)                                                      // all_positions[::n_thin] (:213)
      NQMC_CUDA(cudaMemcpy2DAsync(samples + (size_t)(t / n_thin) * w, (size_t)n_samples * w * sizeof(float), r,
                                  w * sizeof(float), w * sizeof(float), nn, cudaMemcpyDeviceToDevice, st));
  }
  return NQMC_OK;
}

static int energy_core(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, cudaStream_t st) {
  ctx->jas_valid_for = nullptr;
  ctx->act_valid_for = nullptr;
  ctx->act_stored = false;
  if (ctx->sys.kind == NQMC_KIND_SIMPLE) return nq_simple_energy(ctx, r, W, e_l, parts, st);
  if (ctx->sys.use_bf) {
    // tensor-core route: G_i first, then a 5-channel orbital pass carrying the G-weighted Laplacian (nqmc_backflow.cu)
    if (ctx->use_tc && (ctx->sys.H == 64 || ctx->sys.H == 128)) {
      int rc = nq_launch_backflow_xg(ctx, r, W, ctx->x_cur, ctx->gbf, st);
      if (rc) return rc;
      rc = nq_launch_orb_fwd_pc(ctx, ctx->x_cur, W, 5, ctx->phi, ctx->gbf, st);
      if (rc == NQMC_OK) return nq_launch_assemble_bf(ctx, r, W, e_l, parts, 1, st);
      if (rc != NQMC_ERR_UNSUPPORTED) return rc;
    }
    int rc = nq_launch_backflow_xg(ctx, r, W, ctx->x_cur, ctx->gbf, st);
    if (rc) return rc;
    rc = nq_launch_orb_eval(ctx, ctx->x_cur, W, 10, ctx->phi, st);      // FP32 FFMA: value, gradient, Hessian
    if (rc) return rc;
    return nq_launch_assemble_bf(ctx, r, W, e_l, parts, 0, st);
  }
  int rc = nq_launch_orb_eval(ctx, r, W, 5, ctx->phi, st);
  if (rc) return rc;
  return nq_launch_assemble(ctx, r, W, e_l, parts, true, nullptr, nullptr, st);
}

int nqmc_local_energy(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !e_l) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return energy_core(ctx, r, W, e_l, parts, S(st));
}

// reverse pass given that ctx->inv already holds the inverse orbital matrices for r
static int grads_core(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                      double* out, cudaStream_t st) {
  // Plain Slater-Jastrow inside the fused step: orb_bwd_reduce_kernel assigns every network entry and
  // jastrow_from_sums_kernel the four Jastrow scalars -- all P entries -- so the clear is skipped there (it would also sit
  // between the statistics kernel and the reverse-pass kernel and defeat the programmatic dependent launch of the latter).
  const DevSys& sy = ctx->sys;
  const bool all_assigned = sy.kind == NQMC_KIND_ANTISYMMETRIC && !sy.use_cusp && !sy.use_bf && wgt == nullptr &&
                            hdr != nullptr && e_l != nullptr && ctx->jas_valid_for == e_l &&
                            (int64_t)sy.P == (int64_t)sy.n_mlp * ctx->p_mlp + 4;
  if (!all_assigned) NQMC_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)ctx->sys.P, st));
  if (ctx->sys.kind == NQMC_KIND_SIMPLE) return nq_simple_grads(ctx, r, wgt, hdr, e_l, W, out, st);
  int rc = nq_launch_orb_bwd(ctx, ctx->sys.use_bf ? ctx->x_cur : r, wgt, hdr, e_l, W, out, st);
  if (rc) return rc;
  if (wgt == nullptr && hdr != nullptr && e_l != nullptr && ctx->jas_valid_for == e_l)
    rc = nq_launch_jastrow_from_sums(ctx, hdr, out, st);      // sums were taken by the fused phase-A kernel
  else
    rc = nq_launch_walker_grads(ctx, r, wgt, hdr, e_l, W, out, st);
  ctx->jas_valid_for = nullptr;
  if (rc) return rc;
  if (ctx->sys.use_bf) rc = nq_launch_bf_grads(ctx, r, wgt, hdr, e_l, W, out, st);
  return rc;
}

int nqmc_grad_accum(nqmc_ctx* ctx, const float* r, const float* weight, int64_t W, double* out, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !out) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->sys.kind == NQMC_KIND_ANTISYMMETRIC && ctx->sys.use_bf) {
    rc = energy_core(ctx, r, W, nullptr, nullptr, S(st));      // fills x_cur, inv and v (needs orbital gradients)
    if (rc) return rc;
  } else if (ctx->sys.kind == NQMC_KIND_ANTISYMMETRIC) {
    rc = nq_launch_orb_eval(ctx, r, W, 1, ctx->phi, S(st));
    if (rc) return rc;
    rc = nq_launch_assemble_inv(ctx, r, W, S(st));
    if (rc) return rc;
  }
  return grads_core(ctx, r, weight, nullptr, nullptr, W, out, S(st));
}

// O[p][w] = d log|psi(r_w)| / d theta_p for every walker (the O of Stochastic Reconfiguration)
int nqmc_log_psi_grads(nqmc_ctx* ctx, const float* r, int64_t W, float* Ot, int64_t ld, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !Ot || ld < W) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->sys.kind == NQMC_KIND_SIMPLE) return nq_simple_grads_per_walker(ctx, r, W, Ot, ld, S(st));
  if (ctx->sys.use_bf) {
    rc = energy_core(ctx, r, W, nullptr, nullptr, S(st));      // fills x_cur, inv and v
    if (rc) return rc;
  } else {
    rc = nq_launch_orb_eval(ctx, r, W, 1, ctx->phi, S(st));
    if (rc) return rc;
    rc = nq_launch_assemble_inv(ctx, r, W, S(st));
    if (rc) return rc;
  }
  return nq_antisym_grads_per_walker(ctx, r, W, Ot, ld, S(st));
}

int nqmc_energy_stats(nqmc_ctx* ctx, const float* e_l, int64_t W, double* hdr, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!e_l || !hdr) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_launch_stats(ctx, e_l, nullptr, W, hdr, S(st));
}

int nqmc_blocked_stats(nqmc_ctx* ctx, const float* e_l, int64_t n_samples, int64_t W, int32_t n_blocks, double* out3,
                       nqmc_stream st) {
  if (!ctx || !e_l || !out3 || n_samples < 1 || W < 1 || n_blocks < 1 || n_blocks > n_samples) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_blocked_stats(ctx, e_l, n_samples, W, n_blocks, out3, S(st));
}

// Phase A = the MH move (after which the walkers and log|psi|^2 are final) + the energy pass.  The host-buffer entry point
// starts sending the walkers home between the two.
static int step_a_energy(nqmc_ctx* ctx, float* r, int64_t W, float* e_l, double* hdr, nqmc_stream st);

int nqmc_walker_step_a(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms,
                       uint64_t seed, uint64_t step, int64_t wid0, float sigma, int64_t W, float* e_l, float* n_acc,
                       double* hdr, nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !logp || !e_l || !hdr) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  rc = mh_core(ctx, r, logp, normals, uniforms, seed, step, wid0, sigma, W, nullptr, n_acc, S(st));
  if (rc) return rc;
  return step_a_energy(ctx, r, W, e_l, hdr, st);
}

static int step_a_energy(nqmc_ctx* ctx, float* r, int64_t W, float* e_l, double* hdr, nqmc_stream st) {
  int rc;
  ctx->jas_valid_for = nullptr;
  ctx->act_valid_for = nullptr;
  ctx->act_stored = false;
  // phase B of this step runs at the same positions with the same theta: let the 5-channel pass park the activations the
  // width-128 reverse pass needs (NQMC_ACT_REUSE=0 keeps the reverse pass's own value sweep)
  static const bool act_reuse = [] { const char* e = getenv("NQMC_ACT_REUSE"); return !e || atoi(e) != 0; }();
  struct Want { nqmc_ctx* c; ~Want() { c->want_act = false; } } want{ctx};
  ctx->want_act = act_reuse;
  if (ctx->sys.kind == NQMC_KIND_ANTISYMMETRIC && !ctx->sys.use_cusp && !ctx->sys.use_bf &&
      (W + 127) / 128 <= ctx->n_blk_walk) {
    // fused: orbital pass, then E_L assembly + header + Jastrow-scalar sums in one kernel
    rc = nq_launch_orb_eval(ctx, r, W, 5, ctx->phi, S(st));
    if (rc) return rc;
    rc = nq_launch_assemble_stats(ctx, r, W, e_l, ctx->wgt, hdr, S(st));
    if (rc) return rc;
    ctx->jas_valid_for = e_l;
  } else {
    rc = energy_core(ctx, r, W, e_l, nullptr, S(st));
    if (rc) return rc;
    rc = nq_launch_stats(ctx, e_l, ctx->wgt, W, hdr, S(st));
    if (rc) return rc;
  }
  if (ctx->act_stored) { ctx->act_valid_for = e_l; ctx->act_W = W; }
  return NQMC_OK;
}

int nqmc_walker_step_b(nqmc_ctx* ctx, const float* r, const float* e_l, int64_t W, const double* hdr, double* gsum,
                       nqmc_stream st) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r || !e_l || !hdr || !gsum) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return grads_core(ctx, r, nullptr, hdr, e_l, W, gsum, S(st));
}

int nqmc_walker_step(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed,
                     uint64_t step, int64_t wid0, float sigma, int64_t W, float* e_l, float* n_acc, double* hdr,
                     double* gsum, nqmc_stream st) {
  int rc = nqmc_walker_step_a(ctx, r, logp, normals, uniforms, seed, step, wid0, sigma, W, e_l, n_acc, hdr, st);
  if (rc) return rc;
  const bool multi = nq_comm_active(ctx);
  if (multi && (rc = nq_comm_allreduce(ctx, hdr, NQMC_HDR_SIZE, S(st)))) return rc;   // global mean for the centring
  rc = nqmc_walker_step_b(ctx, r, e_l, W, hdr, gsum, st);
  if (rc) return rc;
  if (multi) rc = nq_comm_allreduce(ctx, gsum, ctx->sys.P, S(st));
  return rc;
}

int nqmc_comm_create(nqmc_ctx* ctx, int rank, int world, void* handle_out) {
  if (!ctx) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_comm_create(ctx, rank, world, handle_out);
}

int nqmc_comm_attach(nqmc_ctx* ctx, const void* handles) {
  if (!ctx) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_comm_attach(ctx, handles);
}

int nqmc_comm_active(const nqmc_ctx* ctx) { return ctx && nq_comm_active(ctx) ? 1 : 0; }

int nqmc_comm_allreduce(nqmc_ctx* ctx, double* vec, int64_t len, nqmc_stream st) {
  if (!ctx) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_comm_allreduce(ctx, vec, len, S(st));
}

namespace {
int host_streams(nqmc_ctx* ctx) {
  if (ctx->side_stream) return NQMC_OK;
  NQMC_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  NQMC_CUDA(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
  NQMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_phase_a, cudaEventDisableTiming));
  NQMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_mh_done, cudaEventDisableTiming));
  NQMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_side_done, cudaEventDisableTiming));
  NQMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_upload, cudaEventDisableTiming));
  NQMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_all_done, cudaEventDisableTiming));
  return NQMC_OK;
}

// host inputs of one step -> device staging (AoS -> SoA for the walkers and the normals), on stream st
int stage_inputs(nqmc_ctx* ctx, const float* r_aos, const float* logp, const float* normals_aos, const float* uniforms,
                 int64_t W, float* r_soa, cudaStream_t st) {
  const size_t nn = 3 * (size_t)ctx->sys.N, w = (size_t)W, f = sizeof(float);
  NQMC_CUDA(cudaMemcpyAsync(ctx->aos_stage, r_aos, nn * w * f, cudaMemcpyHostToDevice, st));
  int rc = nq_aos_soa(ctx, ctx->aos_stage, r_soa, W, true, st);
  if (rc) return rc;
  if (normals_aos) {
    NQMC_CUDA(cudaMemcpyAsync(ctx->aos_stage, normals_aos, nn * w * f, cudaMemcpyHostToDevice, st));
    rc = nq_aos_soa(ctx, ctx->aos_stage, ctx->nrm_soa, W, true, st);
    if (rc) return rc;
  }
  if (uniforms) NQMC_CUDA(cudaMemcpyAsync(ctx->uni_stage, uniforms, w * f, cudaMemcpyHostToDevice, st));
  NQMC_CUDA(cudaMemcpyAsync(ctx->logp_stage, logp, w * f, cudaMemcpyHostToDevice, st));
  return NQMC_OK;
}
}  // namespace

int nqmc_walker_step_host_upload(nqmc_ctx* ctx, const float* r_aos, const float* logp, const float* normals_aos,
                                 const float* uniforms, int64_t W) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r_aos || !logp) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  if ((rc = host_streams(ctx))) return rc;
  cudaStream_t up = ctx->up_stream;
  // aos_stage / logp_stage / nrm_soa / uni_stage are free once the phase-A results of the step in flight have gone home
  // (ev_side_done); its reverse pass keeps reading r_soa, which is why the walkers go into the second buffer
  if (ctx->host_step_issued) NQMC_CUDA(cudaStreamWaitEvent(up, ctx->ev_side_done, 0));
  rc = stage_inputs(ctx, r_aos, logp, normals_aos, uniforms, W, ctx->r_soa2, up);
  if (rc) return rc;
  NQMC_CUDA(cudaEventRecord(ctx->ev_upload, up));
  ctx->up_pending = true;
  ctx->up_r = r_aos; ctx->up_logp = logp; ctx->up_nrm = normals_aos; ctx->up_uni = uniforms; ctx->up_W = W;
  return NQMC_OK;
}

int nqmc_walker_step_host_async(nqmc_ctx* ctx, float* r_aos, float* logp, const float* normals_aos, const float* uniforms,
                                uint64_t seed, uint64_t step, int64_t wid0, float sigma, int64_t W, float* e_l, double* hdr,
                                double* gsum, nqmc_stream st_) {
  int rc = check_w(ctx, W);
  if (rc) return rc;
  if (!r_aos || !logp || !e_l || !hdr || !gsum) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  if ((rc = host_streams(ctx))) return rc;
  cudaStream_t st = S(st_);
  const size_t nn = 3 * (size_t)ctx->sys.N, w = (size_t)W, f = sizeof(float);
  bool staged = false;
  if (ctx->up_pending) {
    NQMC_CUDA(cudaStreamWaitEvent(st, ctx->ev_upload, 0));     // also orders an unmatched upload's use of the staging buffers
    staged = ctx->up_r == r_aos && ctx->up_logp == logp && ctx->up_nrm == normals_aos && ctx->up_uni == uniforms && ctx->up_W == W;
    ctx->up_pending = false;
    if (staged) std::swap(ctx->r_soa, ctx->r_soa2);
  }
  if (!staged && (rc = stage_inputs(ctx, r_aos, logp, normals_aos, uniforms, W, ctx->r_soa, st))) return rc;
  const float* nrm = normals_aos ? ctx->nrm_soa : nullptr;
  const float* uni = uniforms ? ctx->uni_stage : nullptr;
  rc = mh_core(ctx, ctx->r_soa, ctx->logp_stage, nrm, uni, seed, step, wid0, sigma, W, nullptr, nullptr, st);
  if (rc) return rc;
  // The walkers and log|psi|^2 are final after the MH move, E_L and the header after the energy pass: they go home on a
  // side stream while the energy pass and the reverse pass run on the caller's stream.  The SoA -> AoS conversion stays on
  // the caller's stream: the persistent tensor-core kernels that follow fill every SM (the reverse pass holds all 227 KB
  // of shared memory, and programmatic dependent launch makes it resident before its predecessor ends), so a conversion
  // kernel on the side stream would only get its turn after them -- only the copy engines run beside them.
  cudaStream_t sd = ctx->side_stream;
  rc = nq_aos_soa(ctx, ctx->r_soa, ctx->aos_stage, W, false, st);
  if (rc) return rc;
  NQMC_CUDA(cudaEventRecord(ctx->ev_mh_done, st));
  NQMC_CUDA(cudaStreamWaitEvent(sd, ctx->ev_mh_done, 0));
  NQMC_CUDA(cudaMemcpyAsync(r_aos, ctx->aos_stage, nn * w * f, cudaMemcpyDeviceToHost, sd));
  NQMC_CUDA(cudaMemcpyAsync(logp, ctx->logp_stage, w * f, cudaMemcpyDeviceToHost, sd));
  rc = step_a_energy(ctx, ctx->r_soa, W, ctx->e_scratch, ctx->hdr_stage, st_);
  if (rc) return rc;
  const bool multi = nq_comm_active(ctx);
  if (multi && (rc = nq_comm_allreduce(ctx, ctx->hdr_stage, NQMC_HDR_SIZE, st))) return rc;
  NQMC_CUDA(cudaEventRecord(ctx->ev_phase_a, st));
  NQMC_CUDA(cudaStreamWaitEvent(sd, ctx->ev_phase_a, 0));
  NQMC_CUDA(cudaMemcpyAsync(e_l, ctx->e_scratch, w * f, cudaMemcpyDeviceToHost, sd));
  NQMC_CUDA(cudaMemcpyAsync(hdr, ctx->hdr_stage, NQMC_HDR_SIZE * sizeof(double), cudaMemcpyDeviceToHost, sd));
  NQMC_CUDA(cudaEventRecord(ctx->ev_side_done, sd));
  ctx->host_step_issued = true;
  rc = nqmc_walker_step_b(ctx, ctx->r_soa, ctx->e_scratch, W, ctx->hdr_stage, ctx->gsum_stage, st_);
  if (rc) return rc;
  if (multi && (rc = nq_comm_allreduce(ctx, ctx->gsum_stage, ctx->sys.P, st))) return rc;
  NQMC_CUDA(cudaMemcpyAsync(gsum, ctx->gsum_stage, (size_t)ctx->sys.P * sizeof(double), cudaMemcpyDeviceToHost, st));
  NQMC_CUDA(cudaStreamWaitEvent(st, ctx->ev_side_done, 0));   // the next call on this stream reuses the staging buffers
  NQMC_CUDA(cudaEventRecord(ctx->ev_all_done, st));
  return NQMC_OK;
}

int nqmc_walker_step_host_wait(nqmc_ctx* ctx, int what) {
  if (!ctx) return NQMC_ERR_ARG;
  if (!ctx->host_step_issued) {
    ctx->err = "nqmc_walker_step_host_wait: no nqmc_walker_step_host_async call to wait for";
    return NQMC_ERR_STATE;
  }
  NQMC_CUDA(cudaSetDevice(ctx->device));
  NQMC_CUDA(cudaEventSynchronize(what == NQMC_WAIT_PHASE_A ? ctx->ev_side_done : ctx->ev_all_done));
  return NQMC_OK;
}

int nqmc_walker_step_host(nqmc_ctx* ctx, float* r_aos, float* logp, const float* normals_aos, const float* uniforms,
                          uint64_t seed, uint64_t step, int64_t wid0, float sigma, int64_t W, float* e_l, double* hdr,
                          double* gsum, nqmc_stream st_) {
  int rc = nqmc_walker_step_host_async(ctx, r_aos, logp, normals_aos, uniforms, seed, step, wid0, sigma, W, e_l, hdr, gsum, st_);
  if (rc) return rc;
  NQMC_CUDA(cudaStreamSynchronize(S(st_)));
  return NQMC_OK;
}

int nqmc_debug_check_guards(nqmc_ctx* ctx) {
  if (!ctx) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  NQMC_CUDA(cudaDeviceSynchronize());
  if (!ctx->guard) return 0;
  int bad = 0;
  std::vector<unsigned char> host(GUARD_BYTES + 256);
  ctx->err.clear();
  for (const auto& g : ctx->guards) {
    if (!g.base) continue;
    bool hit = false;
    NQMC_CUDA(cudaMemcpy(host.data(), g.base, GUARD_BYTES, cudaMemcpyDeviceToHost));               // band in front
    for (size_t i = 0; i < GUARD_BYTES; ++i) hit |= host[i] != GUARD_BYTE;
    const size_t tail = g.padded - g.bytes + GUARD_BYTES;                                           // slack + band behind
    NQMC_CUDA(cudaMemcpy(host.data(), g.base + GUARD_BYTES + g.bytes, tail, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tail; ++i) hit |= host[i] != GUARD_BYTE;
    if (hit) {
      ++bad;
      ctx->err += std::string(ctx->err.empty() ? "guard band overwritten: " : ", ") + g.name;
    }
  }
  return bad;
}

void* nqmc_debug_scratch_ptr(nqmc_ctx* ctx, const char* name) {
  if (!ctx || !name) return nullptr;
  for (const auto& g : ctx->guards)
    if (g.base && std::string(g.name) == name) return g.base + GUARD_BYTES;
  return nullptr;
}

int nqmc_debug_item_iter(int32_t first, int32_t stride, int32_t ns, int64_t n, int32_t* wb, int32_t* e) {
  if (first < 0 || stride < 1 || ns < 1 || n < 0 || !wb || !e) return NQMC_ERR_ARG;
  ItemIter it(first, stride, ns);
  for (int64_t i = 0; i < n; ++i, it.next()) { wb[i] = it.wb; e[i] = it.e; }
  return NQMC_OK;
}

int64_t nqmc_launch_count(const nqmc_ctx* ctx) { return ctx ? ctx->launches : -1; }

int nqmc_use_tensor_cores(nqmc_ctx* ctx, int enable) {
  if (!ctx) return NQMC_ERR_ARG;
  ctx->use_tc = enable != 0;
  return NQMC_OK;
}

int nqmc_profile(nqmc_ctx* ctx, int enable) {
  if (!ctx) return NQMC_ERR_ARG;
  ctx->prof_on = enable != 0;
  for (int k = 0; k < nqmc_ctx::PROF_KINDS; ++k) ctx->prof_n[k] = 0;
  return NQMC_OK;
}

int nqmc_profile_read(nqmc_ctx* ctx, int kind, double* total_ms, int64_t* count) {
  if (!ctx || kind < 0 || kind >= nqmc_ctx::PROF_KINDS || !total_ms || !count) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  NQMC_CUDA(cudaDeviceSynchronize());
  double tot = 0;
  for (int i = 0; i < ctx->prof_n[kind]; ++i) {
    float ms = 0;
    NQMC_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[kind][i][0], ctx->prof_ev[kind][i][1]));
    tot += ms;
  }
  *total_ms = tot;
  *count = ctx->prof_n[kind];
  return NQMC_OK;
}

int nqmc_ffma_peak(nqmc_ctx* ctx, int reps, double* tflops, nqmc_stream st) {
  if (!ctx || !tflops || ctx->cap_w <= 0) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_ffma_peak(ctx, reps, tflops, S(st));
}

int nqmc_rng_fill(nqmc_ctx* ctx, uint64_t seed, uint64_t step, int64_t wid0, int64_t W, float* normals, float* uniforms,
                  nqmc_stream st) {
  if (!ctx || W <= 0) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return nq_rng_fill(ctx, seed, step, wid0, W, normals, uniforms, S(st));
}

}  // extern "C"
