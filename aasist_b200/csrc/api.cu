// C ABI of libaasist_b200.so (declared in include/aasist_b200.h): handle life cycle, strict
// state_dict intake, BN folding / weight packing, and the forward orchestration.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <stdexcept>

#include "common.cuh"
#include "tc.cuh"

namespace aasist {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return AASIST_E_CUDA;
}

int pooled_count(int n, double ratio, int min_nodes) {
  // max(int(n_nodes * k), min) evaluated like the reference's Python expression
  // (models/AASIST.py:315, RawNetGatSpoofST.py:126): IEEE double product, truncation.
  int k = (int)((double)n * ratio);
  return k > min_nodes ? k : min_nodes;
}

// ------------------------------------------------------------------------------------------
// expected state_dict layout
// ------------------------------------------------------------------------------------------
static void expect(aasist_handle* h, const std::string& name, int64_t numel) {
  h->expected.emplace_back(name, numel);
}
static void expect_bn(aasist_handle* h, const std::string& p, int64_t n) {
  expect(h, p + ".weight", n);
  expect(h, p + ".bias", n);
  expect(h, p + ".running_mean", n);
  expect(h, p + ".running_var", n);
}
static void expect_linear(aasist_handle* h, const std::string& p, int64_t in, int64_t out) {
  expect(h, p + ".weight", in * out);
  expect(h, p + ".bias", out);
}
static void expect_encoder(aasist_handle* h, const std::string& enc) {
  for (int i = 0; i < 6; ++i) {
    int ci = h->cfg.enc_channels[i][0], co = h->cfg.enc_channels[i][1];
    std::string p = enc + "." + std::to_string(i) + ".0";
    if (i > 0) expect_bn(h, p + ".bn1", ci);  // dead in forward, present in the checkpoint
    expect(h, p + ".conv1.weight", (int64_t)co * ci * 6);
    expect(h, p + ".conv1.bias", co);
    expect_bn(h, p + ".bn2", co);
    expect(h, p + ".conv2.weight", (int64_t)co * co * 6);
    expect(h, p + ".conv2.bias", co);
    if (ci != co) {
      expect(h, p + ".conv_downsample.weight", (int64_t)co * ci * 3);
      expect(h, p + ".conv_downsample.bias", co);
    }
  }
}
// Res2Net split sizes exactly as Res2NetBlock.__init__ computes them (models/AASIST.py:528-565)
static void res2_splits(int ci, int width, int scale, std::vector<int>& sizes, int& eff_scale) {
  const int w = std::min(width, ci);
  eff_scale = std::min(scale, w);
  const int base = std::max(1, ci / w);
  const int rem = ci - base * (w - 1);
  sizes.clear();
  for (int i = 0; i < w; ++i) sizes.push_back(std::max(1, i < w - 1 ? base : rem));
}
static void expect_encoder_res2(aasist_handle* h, const std::string& enc) {
  for (int i = 0; i < 6; ++i) {
    int ci = h->cfg.enc_channels[i][0], co = h->cfg.enc_channels[i][1];
    std::string p = enc + "." + std::to_string(i) + ".0";
    if (i > 0) expect_bn(h, p + ".bn1", ci);
    std::vector<int> sizes;
    int sc;
    res2_splits(ci, h->cfg.res2net_width, h->cfg.res2net_scale, sizes, sc);
    for (size_t g = 0; g < sizes.size(); ++g) {
      expect(h, p + ".convs." + std::to_string(g) + ".weight", (int64_t)sizes[g] * sizes[g] * 9);
      expect(h, p + ".convs." + std::to_string(g) + ".bias", sizes[g]);
    }
    expect_bn(h, p + ".bn2", ci);
    expect(h, p + ".conv_cat.weight", (int64_t)co * ci * 9);
    expect(h, p + ".conv_cat.bias", co);
    expect(h, p + ".se.fc.0.weight", (int64_t)(co / 16) * co);
    expect(h, p + ".se.fc.2.weight", (int64_t)co * (co / 16));
    if (ci != co) {
      expect(h, p + ".conv_downsample.weight", (int64_t)co * ci * 3);
      expect(h, p + ".conv_downsample.bias", co);
    }
  }
}
static void expect_encoder_33(aasist_handle* h, const std::string& enc) {
  for (int i = 0; i < 6; ++i) {
    int ci = h->cfg.enc_channels[i][0], co = h->cfg.enc_channels[i][1];
    std::string p = enc + "." + std::to_string(i) + ".0";
    if (i > 0) expect_bn(h, p + ".bn1", ci);  // dead in forward (models/AASIST.py:706-712)
    expect(h, p + ".conv1.weight", (int64_t)co * ci * 9);
    expect(h, p + ".conv1.bias", co);
    expect_bn(h, p + ".bn2", co);
    expect(h, p + ".conv2.weight", (int64_t)co * co * 9);
    expect(h, p + ".conv2.bias", co);
    if (ci != co) {
      expect(h, p + ".conv_downsample.weight", (int64_t)co * ci * 3);
      expect(h, p + ".conv_downsample.bias", co);
    }
  }
}
static void expect_gat(aasist_handle* h, const std::string& p, int D, int Do) {
  expect_linear(h, p + ".att_proj", D, Do);
  expect(h, p + ".att_weight", Do);
  expect_linear(h, p + ".proj_with_att", D, Do);
  expect_linear(h, p + ".proj_without_att", D, Do);
  expect_bn(h, p + ".bn", Do);
}
static void expect_htrg(aasist_handle* h, const std::string& p, int D, int Do) {
  expect_linear(h, p + ".proj_type1", D, D);
  expect_linear(h, p + ".proj_type2", D, D);
  expect_linear(h, p + ".att_proj", D, Do);
  expect_linear(h, p + ".att_projM", D, Do);
  expect(h, p + ".att_weight11", Do);
  expect(h, p + ".att_weight22", Do);
  expect(h, p + ".att_weight12", Do);
  expect(h, p + ".att_weightM", Do);
  expect_linear(h, p + ".proj_with_att", D, Do);
  expect_linear(h, p + ".proj_without_att", D, Do);
  expect_linear(h, p + ".proj_with_attM", D, Do);
  expect_linear(h, p + ".proj_without_attM", D, Do);
  expect_bn(h, p + ".bn", Do);
}

static void build_expected(aasist_handle* h) {
  const aasist_config& c = h->cfg;
  const int C = c.enc_channels[5][1];
  if (c.kind == AASIST_KIND_AASIST) {
    const int g0 = c.gat_dims[0], g1 = c.gat_dims[1];
    expect(h, "pos_S", (int64_t)kSpecNodes * C);
    expect(h, "master1", g0);
    expect(h, "master2", g0);
    if (c.spk_emb_dim > 0) {                              // SpeakerConditioningModule (AASIST.py:345-367)
      expect_linear(h, "spk_cond_gat.proj", c.spk_emb_dim, g1);
      if (c.spk_use_attention) {
        expect_linear(h, "spk_cond_gat.attention.0", 2 * g1, g1);
        expect_linear(h, "spk_cond_gat.attention.2", g1, 1);
      }
      expect_linear(h, "spk_cond_gat.fusion.0", 2 * g1, g1);
    }
    expect_bn(h, "first_bn", 1);
    if (c.encoder == AASIST_ENC_RES2NET) expect_encoder_res2(h, "encoder");
    else expect_encoder(h, "encoder");
    expect_gat(h, "GAT_layer_S", C, g0);
    expect_gat(h, "GAT_layer_T", C, g0);
    expect_htrg(h, "HtrgGAT_layer_ST11", g0, g1);
    expect_htrg(h, "HtrgGAT_layer_ST12", g1, g1);
    expect_htrg(h, "HtrgGAT_layer_ST21", g0, g1);
    expect_htrg(h, "HtrgGAT_layer_ST22", g1, g1);
    expect_linear(h, "pool_S.proj", g0, 1);
    expect_linear(h, "pool_T.proj", g0, 1);
    expect_linear(h, "pool_hS1.proj", g1, 1);
    expect_linear(h, "pool_hT1.proj", g1, 1);
    expect_linear(h, "pool_hS2.proj", g1, 1);
    expect_linear(h, "pool_hT2.proj", g1, 1);
    expect_linear(h, "out_layer", 5 * g1, 2);
  } else if (c.kind == AASIST_KIND_ROBUST) {              // models/AASIST_Robust.py:91-196
    const int g0 = c.gat_dims[0], g1 = c.gat_dims[1];
    expect(h, "pos_S", (int64_t)kSpecNodes * C);
    expect(h, "master1", g0);
    expect(h, "master2", g0);                             // defined, never used in forward
    expect(h, "ensemble_weight", 2);
    expect_encoder_33(h, "encoder");
    expect_bn(h, "first_bn", 1);
    expect(h, "gaussian_noise.noise", 1);                 // training-only layers: loaded, never applied in eval
    expect(h, "denoising.g.weight", (int64_t)C * C);
    expect(h, "denoising.g.bias", C);
    expect(h, "denoising.theta.weight", (int64_t)C * C);
    expect(h, "denoising.theta.bias", C);
    expect(h, "denoising.phi.weight", (int64_t)C * C);
    expect(h, "denoising.phi.bias", C);
    expect(h, "denoising.W.weight", (int64_t)C * C);
    expect(h, "denoising.W.bias", C);
    expect_bn(h, "denoising.bn", C);
    expect_gat(h, "GAT_layer_S", C, g0);
    expect_gat(h, "GAT_layer_T", C, g0);
    expect_htrg(h, "HtrgGAT_layer_ST1", g0, g1);
    expect_htrg(h, "HtrgGAT_layer_ST2", g1, g1);
    expect_linear(h, "pool_S.proj", g0, 1);
    expect_linear(h, "pool_T.proj", g0, 1);
    expect_linear(h, "pool_hS.proj", g1, 1);
    expect_linear(h, "pool_hT.proj", g1, 1);
    expect_linear(h, "out_layer", 4 * g1, 2);
    expect_linear(h, "aux_out_layer", C, 2);
  } else {
    expect_bn(h, "first_bn", 1);
    expect_encoder(h, "encoder_T");
    expect_encoder(h, "encoder_S");
    expect_gat(h, "GAT_layer_T", 64, 32);
    expect_gat(h, "GAT_layer_S", 64, 32);
    expect_gat(h, "GAT_layer_ST", 32, 16);
    expect_linear(h, "pool_T.proj", 32, 1);
    expect_linear(h, "pool_S.proj", 32, 1);
    expect_linear(h, "pool_ST.proj", 16, 1);
    expect_linear(h, "proj_T", 14, 12);
    expect_linear(h, "proj_S", 23, 12);
    expect_linear(h, "proj_ST", 16, 1);
    expect_linear(h, "out_layer", 7, 2);
  }
}

// ------------------------------------------------------------------------------------------
// packing helpers (host, double precision folds)
// ------------------------------------------------------------------------------------------
static const std::vector<float>& P(aasist_handle* h, const std::string& name) {
  auto it = h->params.find(name);
  if (it == h->params.end()) throw std::runtime_error("internal: parameter \"" + name + "\" was never declared");
  return it->second;
}

struct BnFold {
  std::vector<double> scale, shift;  // y = scale * x + shift
};
static BnFold fold_bn(aasist_handle* h, const std::string& p) {
  const auto &w = P(h, p + ".weight"), &b = P(h, p + ".bias"), &m = P(h, p + ".running_mean"),
             &v = P(h, p + ".running_var");
  BnFold f;
  f.scale.resize(w.size());
  f.shift.resize(w.size());
  for (size_t i = 0; i < w.size(); ++i) {
    f.scale[i] = (double)w[i] / sqrt((double)v[i] + kBnEps);
    f.shift[i] = (double)b[i] - (double)m[i] * f.scale[i];
  }
  return f;
}

static int upload(float** dst, const std::vector<float>& src) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  AASIST_CUDA(cudaMalloc(dst, sizeof(float) * std::max<size_t>(src.size(), 1)));
  AASIST_CUDA(cudaMemcpy(*dst, src.data(), sizeof(float) * src.size(), cudaMemcpyHostToDevice));
  return 0;
}

static int pack_block_f32(aasist_handle* h, const std::string& p, int ci, int co, ConvBlockF32& blk) {
  const int cop = co <= 32 ? 32 : 64;
  blk.ci = ci;
  blk.co = co;
  blk.downsample = ci != co;
  BnFold bn = fold_bn(h, p + ".bn2");
  const auto &w1 = P(h, p + ".conv1.weight"), &b1 = P(h, p + ".conv1.bias");
  const auto &w2 = P(h, p + ".conv2.weight"), &b2 = P(h, p + ".conv2.bias");
  std::vector<float> pw1((size_t)ci * 6 * cop, 0.f), pb1(cop, 0.f), pw2((size_t)co * 6 * cop, 0.f),
      pb2(cop, 0.f);
  for (int o = 0; o < co; ++o) {
    for (int i = 0; i < ci; ++i)
      for (int t = 0; t < 6; ++t)  // torch layout (co,ci,kh=2,kw=3) -> [ci][kh][kw][cop]
        pw1[((size_t)i * 6 + t) * cop + o] = (float)((double)w1[((size_t)o * ci + i) * 6 + t] * bn.scale[o]);
    pb1[o] = (float)((double)b1[o] * bn.scale[o] + bn.shift[o]);
    for (int i = 0; i < co; ++i)
      for (int t = 0; t < 6; ++t) pw2[((size_t)i * 6 + t) * cop + o] = w2[((size_t)o * co + i) * 6 + t];
    pb2[o] = b2[o];
  }
  int rc;
  if (blk.downsample) {
    const auto &wd = P(h, p + ".conv_downsample.weight"), &bd = P(h, p + ".conv_downsample.bias");
    std::vector<float> pwd((size_t)ci * 3 * cop, 0.f);
    for (int o = 0; o < co; ++o) {
      for (int i = 0; i < ci; ++i)
        for (int t = 0; t < 3; ++t) pwd[((size_t)i * 3 + t) * cop + o] = wd[((size_t)o * ci + i) * 3 + t];
      pb2[o] = (float)((double)b2[o] + (double)bd[o]);
    }
    if ((rc = upload(&blk.wd, pwd))) return rc;
  }
  if ((rc = upload(&blk.w1, pw1))) return rc;
  if ((rc = upload(&blk.b1, pb1))) return rc;
  if ((rc = upload(&blk.w2, pw2))) return rc;
  if ((rc = upload(&blk.b2, pb2))) return rc;
  return 0;
}

static int upload_i(int** dst, const std::vector<int>& src) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  AASIST_CUDA(cudaMalloc(dst, sizeof(int) * std::max<size_t>(src.size(), 1)));
  AASIST_CUDA(cudaMemcpy(*dst, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice));
  return 0;
}

static std::vector<float> bn_pair(const BnFold& f) {       // [2][n]: scale then shift
  std::vector<float> v(2 * f.scale.size());
  for (size_t i = 0; i < f.scale.size(); ++i) {
    v[i] = (float)f.scale[i];
    v[f.scale.size() + i] = (float)f.shift[i];
  }
  return v;
}

// torch conv weight (co,ci,3,3) -> [ci][3][3][cop], optionally scaled per output channel
static std::vector<float> pack_w33(const std::vector<float>& w, int ci, int co, int cop,
                                   const std::vector<double>* scale = nullptr) {
  std::vector<float> t((size_t)ci * 9 * cop, 0.f);
  for (int o = 0; o < co; ++o)
    for (int i = 0; i < ci; ++i)
      for (int k = 0; k < 9; ++k) {
        double v = w[((size_t)o * ci + i) * 9 + k];
        if (scale) v *= (*scale)[o];
        t[((size_t)i * 9 + k) * cop + o] = (float)v;
      }
  return t;
}

static int pack_res2_block(aasist_handle* h, const std::string& p, int index, int ci, int co, Res2BlockF32& blk) {
  const int cop = co <= 32 ? 32 : 64;
  blk.index = index;
  blk.ci = ci;
  blk.co = co;
  blk.first = index == 0;
  blk.downsample = ci != co;
  std::vector<int> sizes;
  int sc;
  res2_splits(ci, h->cfg.res2net_width, h->cfg.res2net_scale, sizes, sc);
  int total = 0;
  for (int n : sizes) total += n;
  if (total != ci) {
    // torch.split(x, split_sizes) raises when the sizes do not add up to the channel count
    set_error("res2net_width=%d does not split %d channels (split sizes sum to %d)", h->cfg.res2net_width, ci, total);
    return AASIST_E_INVALID;
  }
  blk.groups.clear();
  std::vector<float> gw, gb;
  std::vector<int> gw_off;
  int c0 = 0, levels = 1;
  for (size_t g = 0; g < sizes.size(); ++g) {
    Res2Group G;
    G.c0 = c0;
    G.n = sizes[g];
    const bool dep = g > 0 && (int)g % sc == 0;          // sp = sp + spx[i]  (AASIST.py:636-638)
    G.level = dep ? blk.groups[g - 1].level + 1 : 0;
    G.feeds_next = 0;
    if (dep) {
      if (blk.groups[g - 1].n != G.n) {
        set_error("Res2Net split %zu (%d channels) cannot be added to the output of split %zu (%d channels)", g,
                  G.n, g - 1, blk.groups[g - 1].n);
        return AASIST_E_INVALID;                          // reference: RuntimeError on the tensor addition
      }
      blk.groups[g - 1].feeds_next = 1;
    }
    levels = std::max(levels, G.level + 1);
    blk.groups.push_back(G);
    const auto& w = P(h, p + ".convs." + std::to_string(g) + ".weight");
    const auto& b = P(h, p + ".convs." + std::to_string(g) + ".bias");
    gw_off.push_back((int)gw.size());
    gw.insert(gw.end(), w.begin(), w.end());
    gb.insert(gb.end(), b.begin(), b.end());
    c0 += G.n;
  }
  blk.n_groups = (int)blk.groups.size();
  blk.n_levels = levels;
  // launch plan: per dependency level, splits bucketed by the register-accumulator width of the kernel variant
  auto nreg_of = [](int n) {
    const int widths[] = {1, 2, 4, 6, 8, 12, 16, 32, 64};
    for (int w : widths)
      if (n <= w) return w;
    return 64;
  };
  blk.launches.clear();
  std::vector<int> lvl_groups;
  for (int level = 0; level < levels; ++level) {
    const int widths[] = {1, 2, 4, 6, 8, 12, 16, 32, 64};
    for (int w : widths) {
      Res2Launch L{level, w, 0, (int)lvl_groups.size(), 0, 0};
      for (size_t g = 0; g < blk.groups.size(); ++g)
        if (blk.groups[g].level == level && nreg_of(blk.groups[g].n) == w) {
          lvl_groups.push_back((int)g);
          L.n_max = std::max(L.n_max, blk.groups[g].n);
          L.n_tot += blk.groups[g].n;
          ++L.count;
        }
      if (L.count) blk.launches.push_back(L);
    }
  }
  int rc;
  if ((rc = upload_i(&blk.lvl_groups_dev, lvl_groups))) return rc;
  {
    if (blk.groups_dev) cudaFree(blk.groups_dev);
    blk.groups_dev = nullptr;
    AASIST_CUDA(cudaMalloc(&blk.groups_dev, sizeof(Res2Group) * blk.groups.size()));
    AASIST_CUDA(cudaMemcpy(blk.groups_dev, blk.groups.data(), sizeof(Res2Group) * blk.groups.size(),
                           cudaMemcpyHostToDevice));
  }
  if (!blk.first) {
    if ((rc = upload(&blk.bn1, bn_pair(fold_bn(h, p + ".bn1"))))) return rc;
  } else {
    cudaFree(blk.bn1);
    blk.bn1 = nullptr;
  }
  if ((rc = upload(&blk.gw, gw))) return rc;
  if ((rc = upload(&blk.gb, gb))) return rc;
  if ((rc = upload_i(&blk.gw_off, gw_off))) return rc;
  if ((rc = upload(&blk.bn2, bn_pair(fold_bn(h, p + ".bn2"))))) return rc;
  if ((rc = upload(&blk.wcat, pack_w33(P(h, p + ".conv_cat.weight"), ci, co, cop)))) return rc;
  std::vector<float> bcat(cop, 0.f);
  for (int o = 0; o < co; ++o) bcat[o] = P(h, p + ".conv_cat.bias")[o];
  if ((rc = upload(&blk.bcat, bcat))) return rc;
  blk.se_hidden = co / 16;
  if ((rc = upload(&blk.se0, P(h, p + ".se.fc.0.weight")))) return rc;
  if ((rc = upload(&blk.se2, P(h, p + ".se.fc.2.weight")))) return rc;
  if (blk.downsample) {
    const auto &wd = P(h, p + ".conv_downsample.weight"), &bd = P(h, p + ".conv_downsample.bias");
    std::vector<float> pwd((size_t)ci * 3 * cop, 0.f), pbd(cop, 0.f);
    for (int o = 0; o < co; ++o) {
      for (int i = 0; i < ci; ++i)
        for (int t = 0; t < 3; ++t) pwd[((size_t)i * 3 + t) * cop + o] = wd[((size_t)o * ci + i) * 3 + t];
      pbd[o] = bd[o];
    }
    if ((rc = upload(&blk.wd, pwd))) return rc;
    if ((rc = upload(&blk.bd, pbd))) return rc;
  }
  return 0;
}

static int pack_block33(aasist_handle* h, const std::string& p, int ci, int co, ConvBlock33F32& blk) {
  const int cop = co <= 32 ? 32 : 64;
  blk.ci = ci;
  blk.co = co;
  blk.downsample = ci != co;
  BnFold bn = fold_bn(h, p + ".bn2");
  const auto &b1 = P(h, p + ".conv1.bias"), &b2 = P(h, p + ".conv2.bias");
  std::vector<float> pb1(cop, 0.f), pb2(cop, 0.f);
  for (int o = 0; o < co; ++o) {
    pb1[o] = (float)((double)b1[o] * bn.scale[o] + bn.shift[o]);
    pb2[o] = b2[o];
  }
  int rc;
  if (blk.downsample) {
    const auto &wd = P(h, p + ".conv_downsample.weight"), &bd = P(h, p + ".conv_downsample.bias");
    std::vector<float> pwd((size_t)ci * 3 * cop, 0.f);
    for (int o = 0; o < co; ++o) {
      for (int i = 0; i < ci; ++i)
        for (int t = 0; t < 3; ++t) pwd[((size_t)i * 3 + t) * cop + o] = wd[((size_t)o * ci + i) * 3 + t];
      pb2[o] = (float)((double)b2[o] + (double)bd[o]);
    }
    if ((rc = upload(&blk.wd, pwd))) return rc;
  }
  if ((rc = upload(&blk.w1, pack_w33(P(h, p + ".conv1.weight"), ci, co, cop, &bn.scale)))) return rc;
  if ((rc = upload(&blk.b1, pb1))) return rc;
  if ((rc = upload(&blk.w2, pack_w33(P(h, p + ".conv2.weight"), co, co, cop)))) return rc;
  if ((rc = upload(&blk.b2, pb2))) return rc;
  return 0;
}

// graph parameters go into one host vector; device pointers are patched after upload
struct GraphPacker {
  std::vector<float>& buf;
  std::vector<std::pair<const float**, size_t>> fix;
  explicit GraphPacker(std::vector<float>& b) : buf(b) {}
  void put(const float** slot, const std::vector<float>& v) {
    while (buf.size() % 4) buf.push_back(0.f);
    fix.emplace_back(slot, buf.size());
    buf.insert(buf.end(), v.begin(), v.end());
  }
  void patch(const float* base) {
    for (auto& f : fix) *f.first = base + f.second;
  }
};

// torch Linear weight (out,in) -> transposed [in][out], optionally scaled per output
static std::vector<float> transpose_w(const std::vector<float>& w, int in, int out,
                                      const std::vector<double>* scale = nullptr) {
  std::vector<float> t((size_t)in * out);
  for (int o = 0; o < out; ++o)
    for (int i = 0; i < in; ++i)
      t[(size_t)i * out + o] = scale ? (float)((double)w[(size_t)o * in + i] * (*scale)[o]) : w[(size_t)o * in + i];
  return t;
}

static void pack_gat(aasist_handle* h, GraphPacker& gp, const std::string& p, int D, int Do,
                     float temp, GatParams& g) {
  g.D = D;
  g.Do = Do;
  g.temp = temp;
  BnFold bn = fold_bn(h, p + ".bn");
  gp.put(&g.attWt, transpose_w(P(h, p + ".att_proj.weight"), D, Do));
  gp.put(&g.attImg, att_image_floats(P(h, p + ".att_proj.weight"), D, Do));
  gp.put(&g.attB, P(h, p + ".att_proj.bias"));
  gp.put(&g.attW, P(h, p + ".att_weight"));
  gp.put(&g.pWt, transpose_w(P(h, p + ".proj_with_att.weight"), D, Do, &bn.scale));
  gp.put(&g.qWt, transpose_w(P(h, p + ".proj_without_att.weight"), D, Do, &bn.scale));
  std::vector<float> bias(Do);
  const auto &b1 = P(h, p + ".proj_with_att.bias"), &b2 = P(h, p + ".proj_without_att.bias");
  for (int k = 0; k < Do; ++k) bias[k] = (float)(((double)b1[k] + (double)b2[k]) * bn.scale[k] + bn.shift[k]);
  gp.put(&g.bias, bias);
}

static void pack_htrg(aasist_handle* h, GraphPacker& gp, const std::string& p, int D, int Do,
                      float temp, HtrgParams& g) {
  g.D = D;
  g.Do = Do;
  g.temp = temp;
  BnFold bn = fold_bn(h, p + ".bn");
  gp.put(&g.t1Wt, transpose_w(P(h, p + ".proj_type1.weight"), D, D));
  gp.put(&g.t1B, P(h, p + ".proj_type1.bias"));
  gp.put(&g.t2Wt, transpose_w(P(h, p + ".proj_type2.weight"), D, D));
  gp.put(&g.t2B, P(h, p + ".proj_type2.bias"));
  gp.put(&g.attWt, transpose_w(P(h, p + ".att_proj.weight"), D, Do));
  gp.put(&g.attImg, att_image_floats(P(h, p + ".att_proj.weight"), D, Do));
  gp.put(&g.attB, P(h, p + ".att_proj.bias"));
  gp.put(&g.w11, P(h, p + ".att_weight11"));
  gp.put(&g.w22, P(h, p + ".att_weight22"));
  gp.put(&g.w12, P(h, p + ".att_weight12"));
  gp.put(&g.attMWt, transpose_w(P(h, p + ".att_projM.weight"), D, Do));
  gp.put(&g.attMB, P(h, p + ".att_projM.bias"));
  gp.put(&g.wM, P(h, p + ".att_weightM"));
  gp.put(&g.pWt, transpose_w(P(h, p + ".proj_with_att.weight"), D, Do, &bn.scale));
  gp.put(&g.qWt, transpose_w(P(h, p + ".proj_without_att.weight"), D, Do, &bn.scale));
  std::vector<float> bias(Do), biasM(Do);
  const auto &b1 = P(h, p + ".proj_with_att.bias"), &b2 = P(h, p + ".proj_without_att.bias");
  const auto &m1 = P(h, p + ".proj_with_attM.bias"), &m2 = P(h, p + ".proj_without_attM.bias");
  for (int k = 0; k < Do; ++k) {
    bias[k] = (float)(((double)b1[k] + (double)b2[k]) * bn.scale[k] + bn.shift[k]);
    biasM[k] = (float)((double)m1[k] + (double)m2[k]);
  }
  gp.put(&g.bias, bias);
  gp.put(&g.pMWt, transpose_w(P(h, p + ".proj_with_attM.weight"), D, Do));
  gp.put(&g.qMWt, transpose_w(P(h, p + ".proj_without_attM.weight"), D, Do));
  gp.put(&g.biasM, biasM);
}

static void pack_pool(aasist_handle* h, GraphPacker& gp, const std::string& p, int D, PoolParams& g) {
  g.D = D;
  gp.put(&g.w, P(h, p + ".proj.weight"));
  g.b = P(h, p + ".proj.bias")[0];
}

static int pack_graph(aasist_handle* h) {
  const aasist_config& c = h->cfg;
  h->graph_host.clear();
  GraphPacker gp(h->graph_host);
  if (c.kind == AASIST_KIND_AASIST) {
    GraphArgsAasist& a = h->ga;
    memset(&a, 0, sizeof(a));
    a.C = c.enc_channels[5][1];
    a.g0 = c.gat_dims[0];
    a.g1 = c.gat_dims[1];
    gp.put(&a.posS, P(h, "pos_S"));
    gp.put(&a.master1, P(h, "master1"));
    gp.put(&a.master2, P(h, "master2"));
    pack_gat(h, gp, "GAT_layer_S", a.C, a.g0, (float)c.temperatures[0], a.gatS);
    pack_gat(h, gp, "GAT_layer_T", a.C, a.g0, (float)c.temperatures[1], a.gatT);
    // all four heterogeneous layers use temperatures[2] (models/AASIST.py:785-794)
    pack_htrg(h, gp, "HtrgGAT_layer_ST11", a.g0, a.g1, (float)c.temperatures[2], a.st11);
    pack_htrg(h, gp, "HtrgGAT_layer_ST12", a.g1, a.g1, (float)c.temperatures[2], a.st12);
    pack_htrg(h, gp, "HtrgGAT_layer_ST21", a.g0, a.g1, (float)c.temperatures[2], a.st21);
    pack_htrg(h, gp, "HtrgGAT_layer_ST22", a.g1, a.g1, (float)c.temperatures[2], a.st22);
    pack_pool(h, gp, "pool_S", a.g0, a.poolS);
    pack_pool(h, gp, "pool_T", a.g0, a.poolT);
    pack_pool(h, gp, "pool_hS1", a.g1, a.poolhS1);
    pack_pool(h, gp, "pool_hT1", a.g1, a.poolhT1);
    pack_pool(h, gp, "pool_hS2", a.g1, a.poolhS2);
    pack_pool(h, gp, "pool_hT2", a.g1, a.poolhT2);
    gp.put(&a.outWt, transpose_w(P(h, "out_layer.weight"), 5 * a.g1, 2));
    a.outB0 = P(h, "out_layer.bias")[0];
    a.outB1 = P(h, "out_layer.bias")[1];
    if (c.spk_emb_dim > 0) {
      SpkParams& sp = a.spk;
      sp.emb_dim = c.spk_emb_dim;
      sp.use_attention = c.spk_use_attention ? 1 : 0;
      gp.put(&sp.projW, P(h, "spk_cond_gat.proj.weight"));
      gp.put(&sp.projB, P(h, "spk_cond_gat.proj.bias"));
      if (sp.use_attention) {
        gp.put(&sp.att0Wt, transpose_w(P(h, "spk_cond_gat.attention.0.weight"), 2 * a.g1, a.g1));
        gp.put(&sp.att0B, P(h, "spk_cond_gat.attention.0.bias"));
        gp.put(&sp.att2W, P(h, "spk_cond_gat.attention.2.weight"));
        sp.att2B = P(h, "spk_cond_gat.attention.2.bias")[0];
      }
      gp.put(&sp.fusWt, transpose_w(P(h, "spk_cond_gat.fusion.0.weight"), 2 * a.g1, a.g1));
      gp.put(&sp.fusB, P(h, "spk_cond_gat.fusion.0.bias"));
    }
  } else if (c.kind == AASIST_KIND_ROBUST) {
    GraphArgsAasist& a = h->ga;
    memset(&a, 0, sizeof(a));
    a.robust = 1;
    a.C = c.enc_channels[5][1];
    a.g0 = c.gat_dims[0];
    a.g1 = c.gat_dims[1];
    gp.put(&a.posS, P(h, "pos_S"));
    gp.put(&a.master1, P(h, "master1"));
    a.master2 = nullptr;
    pack_gat(h, gp, "GAT_layer_S", a.C, a.g0, (float)c.temperatures[0], a.gatS);
    pack_gat(h, gp, "GAT_layer_T", a.C, a.g0, (float)c.temperatures[1], a.gatT);
    pack_htrg(h, gp, "HtrgGAT_layer_ST1", a.g0, a.g1, (float)c.temperatures[2], a.st11);   // AASIST_Robust.py:146-156
    pack_htrg(h, gp, "HtrgGAT_layer_ST2", a.g1, a.g1, (float)c.temperatures[3], a.st12);
    pack_pool(h, gp, "pool_S", a.g0, a.poolS);
    pack_pool(h, gp, "pool_T", a.g0, a.poolT);
    pack_pool(h, gp, "pool_hS", a.g1, a.poolhS1);
    pack_pool(h, gp, "pool_hT", a.g1, a.poolhT1);
    gp.put(&a.outWt, transpose_w(P(h, "out_layer.weight"), 4 * a.g1, 2));
    a.outB0 = P(h, "out_layer.bias")[0];
    a.outB1 = P(h, "out_layer.bias")[1];
    gp.put(&a.auxWt, transpose_w(P(h, "aux_out_layer.weight"), a.C, 2));
    a.auxB0 = P(h, "aux_out_layer.bias")[0];
    a.auxB1 = P(h, "aux_out_layer.bias")[1];
    // F.softmax(ensemble_weight, dim=0) (AASIST_Robust.py:293), evaluated like torch: exp(x - max) / sum in fp32
    const float e0 = P(h, "ensemble_weight")[0], e1 = P(h, "ensemble_weight")[1];
    const float mx = std::max(e0, e1);
    const float x0 = expf(e0 - mx), x1 = expf(e1 - mx);
    a.ens0 = x0 / (x0 + x1);
    a.ens1 = x1 / (x0 + x1);
  } else {
    GraphArgsRawGat& a = h->gr;
    memset(&a, 0, sizeof(a));
    pack_gat(h, gp, "GAT_layer_T", 64, 32, 1.f, a.gatT);
    pack_gat(h, gp, "GAT_layer_S", 64, 32, 1.f, a.gatS);
    pack_gat(h, gp, "GAT_layer_ST", 32, 16, 1.f, a.gatST);
    pack_pool(h, gp, "pool_T", 32, a.poolT);
    pack_pool(h, gp, "pool_S", 32, a.poolS);
    pack_pool(h, gp, "pool_ST", 16, a.poolST);
    gp.put(&a.projTW, P(h, "proj_T.weight"));
    gp.put(&a.projTB, P(h, "proj_T.bias"));
    gp.put(&a.projSW, P(h, "proj_S.weight"));
    gp.put(&a.projSB, P(h, "proj_S.bias"));
    gp.put(&a.projSTW, P(h, "proj_ST.weight"));
    a.projSTB = P(h, "proj_ST.bias")[0];
    gp.put(&a.outW, P(h, "out_layer.weight"));
    gp.put(&a.outB, P(h, "out_layer.bias"));
  }
  int rc = upload(&h->graph_buf, h->graph_host);
  if (rc) return rc;
  gp.patch(h->graph_buf);
  return 0;
}

// Makes the handle's device current for the duration of an entry point and restores the caller's
// current device on exit (a handle on cuda:1 must not silently switch the thread to cuda:1).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int enter(aasist_handle* h) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice (no usable CUDA device; there is no CPU fallback)");
    if (h && h->device < 0) h->device = prev;
    if (h && prev != h->device) {
      AASIST_CUDA(cudaSetDevice(h->device));
      switched = true;
    }
    return 0;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define AASIST_ENTER_DEVICE(h)        \
  DeviceGuard _device_guard;          \
  int rc = _device_guard.enter(h);    \
  if (rc) return rc

// pooled width of the sinc front end: floor(frames / 3), frames = floor((L - taps) / stride) + 1
static inline int front_width(const aasist_handle* h, int L) {
  if (L < h->taps) return 0;
  return ((L - h->taps) / h->stride + 1) / 3;
}

struct Plan {          // activation sizes per utterance
  int W[7];            // W[0] = frontend width, W[i+1] = width after block i
  size_t front, act, mid, enc;
};
static int make_plan(const aasist_handle* h, int L, Plan& pl) {
  pl.W[0] = front_width(h, L);
  if (L < h->taps || pl.W[0] < 1) {
    set_error("input of %d samples is shorter than the %d-tap sinc filters", L, h->taps);
    return AASIST_E_INVALID;
  }
  for (int i = 0; i < 6; ++i) {
    pl.W[i + 1] = pl.W[i] / 3;
    if (pl.W[i + 1] < 1) {
      // reference: RuntimeError from max_pool2d ("Output size is too small") for L < 2315
      set_error("input of %d samples is too short: encoder block %d would pool %d columns to 0 "
                "(Output size is too small; need %d pooled front-end columns)", L, i, pl.W[i], 729);
      return AASIST_E_INVALID;
    }
  }
  size_t act = 0, mid = 0;
  for (int i = 0; i < 6; ++i) {
    int co = h->cfg.enc_channels[i][1];
    act = std::max(act, (size_t)co * kSpecNodes * pl.W[i + 1]);
    mid = std::max(mid, (size_t)co * 24 * pl.W[i]);
  }
  pl.front = (size_t)kSpecNodes * pl.W[0];
  pl.act = act;
  pl.mid = mid;
  pl.enc = (size_t)h->cfg.enc_channels[5][1] * kSpecNodes * pl.W[6];
  return 0;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
constexpr int kChunkF32 = 32;  // utterances per encoder pass on the fp32 path (bounds scratch)

// does the whole encoder run on the tensor-core path?
static inline bool tc_encoder(const aasist_handle* h) {
  return h->cfg.precision != AASIST_PREC_FP32 && h->cfg.encoder == AASIST_ENC_RESIDUAL23 &&
         h->cfg.kind != AASIST_KIND_ROBUST;
}
// tensor-core sinc front end feeding fp32 encoder kernels (Res2Net encoder with precision f16x3)
static inline bool tc_frontend_only(const aasist_handle* h) {
  return h->cfg.precision != AASIST_PREC_FP32 && h->cfg.encoder != AASIST_ENC_RESIDUAL23 &&
         h->cfg.kind != AASIST_KIND_ROBUST;
}

// fp32 scratch (floats) one encoder block needs beyond its input/output, for nb utterances
static size_t block_scratch_floats(const aasist_handle* h, const Plan& pl, int nb) {
  size_t m = 0;
  for (int i = 0; i < 6; ++i) {
    if (h->cfg.encoder == AASIST_ENC_RES2NET) m = std::max(m, res2_block_scratch_floats(h->res2[i], nb, pl.W[i]));
    else m = std::max(m, (size_t)nb * h->cfg.enc_channels[i][1] * 24 * pl.W[i]);
  }
  return m;
}

static int ensure_own_ws(aasist_handle* h, size_t bytes) {
  if (h->own_ws_bytes >= bytes) return 0;
  cudaFree(h->own_ws);                 // synchronises with any work still using the old buffer
  h->own_ws = nullptr;
  h->own_ws_bytes = 0;
  AASIST_CUDA(cudaMalloc(&h->own_ws, bytes));
  h->own_ws_bytes = bytes;
  return 0;
}

}  // namespace aasist

using namespace aasist;

// ==========================================================================================
// C ABI
// ==========================================================================================
#pragma GCC visibility push(default)
extern "C" {

int aasist_abi_version(void) { return AASIST_B200_ABI_VERSION; }
const char* aasist_last_error(void) { return g_err; }

int aasist_create(const aasist_config* cfg, aasist_handle** out) {
  if (!cfg || !out) {
    set_error("aasist_create: null argument");
    return AASIST_E_INVALID;
  }
  *out = nullptr;
  if (cfg->kind != AASIST_KIND_AASIST && cfg->kind != AASIST_KIND_RAWGAT_ST && cfg->kind != AASIST_KIND_ROBUST) {
    set_error("unknown model kind %d", cfg->kind);
    return AASIST_E_INVALID;
  }
  if (cfg->precision != AASIST_PREC_FP32 && cfg->precision != AASIST_PREC_F16X3 &&
      cfg->precision != AASIST_PREC_F16X2) {
    set_error("unknown precision mode %d", cfg->precision);
    return AASIST_E_INVALID;
  }
  if (cfg->first_conv < 3 || cfg->first_conv > 1025) {
    set_error("first_conv=%d out of range", cfg->first_conv);
    return AASIST_E_INVALID;
  }
  const bool robust = cfg->kind == AASIST_KIND_ROBUST;
  if (!robust && cfg->n_filters / 3 != kSpecNodes) {
    // pos_S is (1,23,C) (models/AASIST.py:774): filts[0] must pool (3x) to 23 bands.  The Robust model builds
    // whatever `first_conv` says and only fails in forward (AASIST_Robust.py:237-238): checked there.
    set_error("filts[0]=%d must give 23 pooled bands (69..71)", cfg->n_filters);
    return AASIST_E_INVALID;
  }
  if (robust && (cfg->n_filters < 3 || cfg->n_filters > 256)) {
    set_error("AASIST-Robust: %d sinc filters out of range (3..256)", cfg->n_filters);
    return AASIST_E_INVALID;
  }
  const int enc_kind = cfg->encoder;
  if ((robust && enc_kind != AASIST_ENC_RESIDUAL33) ||
      (cfg->kind == AASIST_KIND_RAWGAT_ST && enc_kind != AASIST_ENC_RESIDUAL23) ||
      (cfg->kind == AASIST_KIND_AASIST && enc_kind != AASIST_ENC_RESIDUAL23 && enc_kind != AASIST_ENC_RES2NET)) {
    set_error("encoder type %d is not valid for model kind %d", enc_kind, cfg->kind);
    return AASIST_E_INVALID;
  }
  if (enc_kind == AASIST_ENC_RES2NET && (cfg->res2net_width < 1 || cfg->res2net_scale < 1)) {
    set_error("res2net_width=%d / res2net_scale=%d must be >= 1", cfg->res2net_width, cfg->res2net_scale);
    return AASIST_E_INVALID;
  }
  for (int i = 0; i < 6; ++i) {
    int ci = cfg->enc_channels[i][0], co = cfg->enc_channels[i][1];
    int prev = i == 0 ? 1 : cfg->enc_channels[i - 1][1];
    if (ci != prev || co < 1 || co > 64) {
      set_error("encoder block %d channels (%d,%d) invalid (input must be %d, output 1..64)", i, ci, co, prev);
      return AASIST_E_INVALID;
    }
  }
  if (cfg->kind != AASIST_KIND_RAWGAT_ST) {
    if (cfg->gat_dims[0] < 1 || cfg->gat_dims[0] > 64 || cfg->gat_dims[1] < 1 || cfg->gat_dims[1] > 64) {
      set_error("gat_dims (%d,%d) must be within 1..64", cfg->gat_dims[0], cfg->gat_dims[1]);
      return AASIST_E_INVALID;
    }
    for (int i = 0; i < (robust ? 4 : 3); ++i)
      if (!(cfg->pool_ratios[i] > 0.0) || !(cfg->temperatures[i] != 0.0)) {
        set_error("pool_ratios[%d] / temperatures[%d] invalid", i, i);
        return AASIST_E_INVALID;
      }
    if (cfg->spk_emb_dim < 0 || cfg->spk_emb_dim > 4096 || (robust && cfg->spk_emb_dim != 0)) {
      set_error("spk_emb_dim=%d invalid", cfg->spk_emb_dim);
      return AASIST_E_INVALID;
    }
  } else if (cfg->enc_channels[5][1] != 64) {
    set_error("RawGAT-ST requires 64 encoder output channels (GraphAttentionLayer(64,32))");
    return AASIST_E_INVALID;
  }
  aasist_handle* h = new aasist_handle();
  h->cfg = *cfg;
  if (h->cfg.sample_rate <= 0) h->cfg.sample_rate = 16000;
  h->taps = cfg->first_conv % 2 == 0 ? cfg->first_conv + 1 : cfg->first_conv;  // AASIST.py:449-450
  h->stride = robust ? 256 : 1;                                                 // AASIST_Robust.py:100
  h->n_encoders = cfg->kind == AASIST_KIND_RAWGAT_ST ? 2 : 1;
  // the handle can be created and fed parameters without a device (host-side checks only);
  // every compute entry point, starting with aasist_finalize, requires one.
  if (cudaGetDevice(&h->device) != cudaSuccess) {
    cudaGetLastError();
    h->device = -1;
  }
  build_expected(h);
  *out = h;
  return AASIST_OK;
}

static void score_release(aasist_handle* h) {
  aasist_handle::ScoreStream& sc = h->score;
  for (int i = 0; i < 2; ++i) {
    if (sc.pin[i]) cudaFreeHost(sc.pin[i]);
    cudaFree(sc.dx[i]);
    if (sc.h2d_done[i]) cudaEventDestroy(sc.h2d_done[i]);
    if (sc.fwd_done[i]) cudaEventDestroy(sc.fwd_done[i]);
    sc.pin[i] = sc.dx[i] = nullptr;
    sc.h2d_done[i] = sc.fwd_done[i] = nullptr;
  }
  cudaFree(sc.d_logits);
  cudaFree(sc.d_hidden);
  sc = aasist_handle::ScoreStream();
}

int aasist_destroy(aasist_handle* h) {
  if (!h) return AASIST_OK;
  DeviceGuard guard;
  if (h->device >= 0) guard.enter(h);
  cudaFree(h->bank);
  cudaFree(h->bank_t);
  for (int e = 0; e < 2; ++e)
    for (int i = 0; i < 6; ++i) {
      ConvBlockF32& b = h->blocks[e][i];
      cudaFree(b.w1); cudaFree(b.b1); cudaFree(b.w2); cudaFree(b.b2); cudaFree(b.wd);
    }
  for (int i = 0; i < 6; ++i) {
    Res2BlockF32& r = h->res2[i];
    cudaFree(r.groups_dev); cudaFree(r.lvl_groups_dev); cudaFree(r.bn1); cudaFree(r.gw); cudaFree(r.gb); cudaFree(r.gw_off); cudaFree(r.bn2);
    cudaFree(r.wcat); cudaFree(r.bcat); cudaFree(r.se0); cudaFree(r.se2); cudaFree(r.wd); cudaFree(r.bd);
    ConvBlock33F32& b = h->blocks33[i];
    cudaFree(b.w1); cudaFree(b.b1); cudaFree(b.w2); cudaFree(b.b2); cudaFree(b.wd);
  }
  cudaFree(h->graph_buf);
  tc_destroy(h);
  for (auto& sp : h->prof_pending) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : h->prof_pool) cudaEventDestroy(e);
  if (h->pin_x) cudaFreeHost(h->pin_x);
  if (h->pin_out) cudaFreeHost(h->pin_out);
  if (h->copy_stream) {
    cudaStreamDestroy(h->copy_stream);
    cudaEventDestroy(h->copy_done[0]); cudaEventDestroy(h->copy_done[1]); cudaEventDestroy(h->start_ev);
  }
  score_release(h);
  cudaFree(h->dev_stage);
  cudaFree(h->stage_meta);
  cudaFree(h->own_ws);
  cudaFree(h->front_bimg_masked);
  if (h->range_flag) cudaFreeHost(h->range_flag);
  delete h;
  return AASIST_OK;
}

int aasist_num_params(const aasist_handle* h) { return h ? (int)h->expected.size() : 0; }

const char* aasist_param_name(const aasist_handle* h, int index, int64_t* numel) {
  if (!h || index < 0 || index >= (int)h->expected.size()) return nullptr;
  if (numel) *numel = h->expected[index].second;
  return h->expected[index].first.c_str();
}

int aasist_set_param(aasist_handle* h, const char* name, const float* data, int64_t numel) {
  if (!h || !name || (!data && numel > 0)) {
    set_error("aasist_set_param: null argument");
    return AASIST_E_INVALID;
  }
  std::string n(name);
  const std::string nbt = "num_batches_tracked";
  if (n.size() >= nbt.size() && n.compare(n.size() - nbt.size(), nbt.size(), nbt) == 0) return AASIST_OK;
  auto it = std::find_if(h->expected.begin(), h->expected.end(),
                         [&](const std::pair<std::string, int64_t>& p) { return p.first == n; });
  if (it == h->expected.end()) {
    set_error("unexpected key in state_dict: \"%s\"", name);
    return AASIST_E_PARAM;
  }
  if (it->second != numel) {
    set_error("size mismatch for %s: expected %lld elements, got %lld", name, (long long)it->second,
              (long long)numel);
    return AASIST_E_PARAM;
  }
  std::vector<float> v((size_t)numel);
  cudaPointerAttributes attr;
  bool on_device = false;
  if (numel > 0 && cudaPointerGetAttributes(&attr, data) == cudaSuccess)
    on_device = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
  else
    cudaGetLastError();
  if (on_device) {
    AASIST_CUDA(cudaMemcpy(v.data(), data, sizeof(float) * numel, cudaMemcpyDeviceToHost));
  } else if (numel > 0) {
    memcpy(v.data(), data, sizeof(float) * numel);
  }
  h->params[n] = std::move(v);
  h->finalized = false;
  return AASIST_OK;
}

static int finalize_impl(aasist_handle* h);
int aasist_finalize(aasist_handle* h) {
  if (!h) return AASIST_E_INVALID;
  try {
    return finalize_impl(h);
  } catch (const std::exception& e) {            // no exception crosses the C ABI
    set_error("aasist_finalize: %s", e.what());
    return AASIST_E_PARAM;
  }
}
static int finalize_impl(aasist_handle* h) {
  AASIST_ENTER_DEVICE(h);
  for (auto& p : h->expected)
    if (!h->params.count(p.first)) {
      set_error("missing key in state_dict: \"%s\"", p.first.c_str());
      return AASIST_E_PARAM;
    }
  if ((rc = build_filterbank(h))) return rc;
  BnFold bn0 = fold_bn(h, "first_bn");
  h->bn0_scale = (float)bn0.scale[0];
  h->bn0_shift = (float)bn0.shift[0];
  const char* enc_names[2] = {h->cfg.kind == AASIST_KIND_RAWGAT_ST ? "encoder_T" : "encoder", "encoder_S"};
  for (int e = 0; e < h->n_encoders; ++e)
    for (int i = 0; i < 6; ++i) {
      std::string p = std::string(enc_names[e]) + "." + std::to_string(i) + ".0";
      const int ci = h->cfg.enc_channels[i][0], co = h->cfg.enc_channels[i][1];
      if (h->cfg.encoder == AASIST_ENC_RES2NET) rc = pack_res2_block(h, p, i, ci, co, h->res2[i]);
      else if (h->cfg.encoder == AASIST_ENC_RESIDUAL33) rc = pack_block33(h, p, ci, co, h->blocks33[i]);
      else rc = pack_block_f32(h, p, ci, co, h->blocks[e][i]);
      if (rc) return rc;
    }
  if ((rc = pack_graph(h))) return rc;
  if (h->cfg.precision != AASIST_PREC_FP32 && h->cfg.kind != AASIST_KIND_ROBUST)
    if ((rc = tc_finalize(h))) return rc;
  h->finalized = true;
  return AASIST_OK;
}

int aasist_hidden_dim(const aasist_handle* h) {
  if (!h) return 0;
  if (h->cfg.kind == AASIST_KIND_ROBUST) return 2;        // (ensemble_logits, logits), AASIST_Robust.py:303
  return h->cfg.kind == AASIST_KIND_AASIST ? 5 * h->cfg.gat_dims[1] : 7;
}

int aasist_topk_layout(const aasist_handle* h, int32_t L, int32_t* n_pools, int32_t* nk) {
  if (!h) return AASIST_E_INVALID;
  Plan pl;
  int rc = make_plan(h, L, pl);
  if (rc) return rc;
  const int NT = pl.W[6];
  int tmp[12];
  int np;
  if (h->cfg.kind == AASIST_KIND_AASIST) {
    const double* r = h->cfg.pool_ratios;
    int nS = pooled_count(kSpecNodes, r[0], 1), nT = pooled_count(NT, r[1], 1);
    int nS2 = pooled_count(nS, r[2], 1), nT2 = pooled_count(nT, r[2], 1);
    int v[12] = {kSpecNodes, nS, NT, nT, nS, nS2, nT, nT2, nS, nS2, nT, nT2};
    memcpy(tmp, v, sizeof(v));
    np = 6;
  } else if (h->cfg.kind == AASIST_KIND_ROBUST) {           // pool_S, pool_T, pool_hS, pool_hT
    const double* r = h->cfg.pool_ratios;
    int nS = pooled_count(kSpecNodes, r[0], 1), nT = pooled_count(NT, r[1], 1);
    int v[8] = {kSpecNodes, nS, NT, nT, nS, pooled_count(nS, r[2], 1), nT, pooled_count(nT, r[3], 1)};
    memcpy(tmp, v, sizeof(v));
    np = 4;
  } else {
    int v[6] = {kSpecNodes, pooled_count(kSpecNodes, 0.64, 2), NT, pooled_count(NT, 0.81, 2), 12,
                pooled_count(12, 0.64, 2)};
    memcpy(tmp, v, sizeof(v));
    np = 3;
  }
  if (n_pools) *n_pools = np;
  if (nk) memcpy(nk, tmp, sizeof(int) * 2 * np);
  int total = 0;
  for (int i = 0; i < np; ++i) total += tmp[2 * i + 1];
  return total;
}

int64_t aasist_workspace_bytes(const aasist_handle* h, int32_t B, int32_t L) {
  if (!h || B < 1) {
    set_error("aasist_workspace_bytes: invalid arguments");
    return AASIST_E_INVALID;
  }
  Plan pl;
  int rc = make_plan(h, L, pl);
  if (rc) return rc;
  size_t enc_all = align256(sizeof(float) * pl.enc * B) * h->n_encoders;
  if (tc_encoder(h)) return (int64_t)(enc_all + tc_workspace_bytes(h, B, L));
  int nb = std::min<int>(B, kChunkF32);
  size_t bytes = enc_all + align256(sizeof(float) * pl.front * nb) +
                 2 * align256(sizeof(float) * pl.act * nb) + align256(sizeof(float) * block_scratch_floats(h, pl, nb));
  return (int64_t)bytes;
}

static int run_encoder_f32(aasist_handle* h, int enc, const float* front, int nb, const Plan& pl,
                           float* actA, float* actB, float* scratch, float* enc_out, cudaStream_t st) {
  const float* in = front;
  for (int i = 0; i < 6; ++i) {
    float* out = i == 5 ? enc_out : (i % 2 == 0 ? actA : actB);
    int rc;
    if (h->cfg.encoder == AASIST_ENC_RES2NET) rc = launch_res2_block(h, h->res2[i], in, nb, pl.W[i], scratch, out, st);
    else if (h->cfg.encoder == AASIST_ENC_RESIDUAL33)
      rc = launch_block33_f32(h, h->blocks33[i], in, nb, pl.W[i], scratch, out, st);
    else rc = launch_block_f32(h, h->blocks[enc][i], in, nb, pl.W[i], scratch, out, st);
    if (rc) return rc;
    in = out;
  }
  return 0;
}

int aasist_forward_ex(aasist_handle* h, const float* x, int32_t B, int32_t L, const aasist_forward_opts* opts,
                      float* last_hidden, float* logits, int32_t* topk_idx, float* pool_scores, void* workspace,
                      int64_t workspace_bytes, void* stream) {
  if (!h || !x || !last_hidden || !logits || B < 1) {
    set_error("aasist_forward: invalid arguments");
    return AASIST_E_INVALID;
  }
  if (!h->finalized) {
    set_error("aasist_forward called before aasist_finalize");
    return AASIST_E_STATE;
  }
  AASIST_ENTER_DEVICE(h);
  const aasist_config& c = h->cfg;
  if (c.kind == AASIST_KIND_ROBUST && c.n_filters / 3 != kSpecNodes) {
    // reference: `e_S.transpose(1, 2) + self.pos_S` fails for every input (AASIST_Robust.py:237-238)
    set_error("The size of tensor a (%d) must match the size of tensor b (23) at non-singleton dimension 1: "
              "first_conv=%d gives %d spectral bands but pos_S is (1,23,C)", c.n_filters / 3, c.n_filters,
              c.n_filters / 3);
    return AASIST_E_INVALID;
  }
  Plan pl;
  if ((rc = make_plan(h, L, pl))) return rc;
  int mask_start = 0, mask_count = 0;
  const float* spk = nullptr;
  if (opts) {
    if (opts->freq_mask_count < 0 || opts->freq_mask_start < 0 ||
        opts->freq_mask_start + opts->freq_mask_count > c.n_filters) {
      set_error("Freq_aug mask rows [%d, %d) outside the %d-filter bank", opts->freq_mask_start,
                opts->freq_mask_start + opts->freq_mask_count, c.n_filters);
      return AASIST_E_INVALID;
    }
    mask_start = opts->freq_mask_start;
    mask_count = opts->freq_mask_count;
    // `if self.use_speaker_conditioning and speaker_embedding is not None` (AASIST.py:895): an embedding passed
    // to a model without the module is ignored, like in the reference
    if (opts->speaker_embedding && c.kind == AASIST_KIND_AASIST && c.spk_emb_dim > 0) {
      if (c.spk_level != 0) {
        // reference: the utterance-level branch applies fusion Linear(2*g1, g1) to cat(last_hidden (5*g1),
        // proj (g1)) and raises (AASIST.py:913-916, :412)
        set_error("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d): utterance-level speaker conditioning "
                  "is not runnable in the reference either", B, 6 * c.gat_dims[1], 2 * c.gat_dims[1], c.gat_dims[1]);
        return AASIST_E_INVALID;
      }
      spk = opts->speaker_embedding;
    }
  }
  int64_t need = aasist_workspace_bytes(h, B, L);
  if (need < 0) return (int)need;
  if (!workspace) {
    if ((rc = ensure_own_ws(h, (size_t)need))) return rc;
    workspace = h->own_ws;
    workspace_bytes = (int64_t)h->own_ws_bytes;
  }
  if (workspace_bytes < need) {
    set_error("workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
    return AASIST_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* enc_out[2];
  size_t enc_bytes = align256(sizeof(float) * pl.enc * B);
  for (int e = 0; e < h->n_encoders; ++e) enc_out[e] = (float*)(ws + e * enc_bytes);
  ws += enc_bytes * h->n_encoders;
  if (tc_encoder(h)) {
    if ((rc = tc_encode(h, x, B, L, enc_out, ws, mask_start, mask_count, st))) return rc;
  } else {
    int nbmax = std::min<int>(B, kChunkF32);
    float* front = (float*)ws;
    ws += align256(sizeof(float) * pl.front * nbmax);
    float* actA = (float*)ws;
    ws += align256(sizeof(float) * pl.act * nbmax);
    float* actB = (float*)ws;
    ws += align256(sizeof(float) * pl.act * nbmax);
    float* scratch = (float*)ws;
    for (int b0 = 0; b0 < B; b0 += nbmax) {
      int nb = std::min(nbmax, B - b0);
      const float* xb = x + (size_t)b0 * L;
      if (c.kind == AASIST_KIND_ROBUST) rc = launch_frontend_strided_f32(h, xb, nb, L, front, mask_start, mask_count, st);
      else if (tc_frontend_only(h)) rc = tc_frontend_to_f32(h, xb, nb, L, front, mask_start, mask_count, st);
      else rc = launch_frontend_f32(h, xb, nb, L, front, mask_start, mask_count, st);
      if (rc) return rc;
      for (int e = 0; e < h->n_encoders; ++e)
        if ((rc = run_encoder_f32(h, e, front, nb, pl, actA, actB, scratch, enc_out[e] + (size_t)b0 * pl.enc, st)))
          return rc;
    }
  }
  if (c.kind != AASIST_KIND_RAWGAT_ST)
    return launch_graph_aasist(h, enc_out[0], B, pl.W[6], spk, last_hidden, logits, topk_idx, pool_scores, st);
  return launch_graph_rawgat(h, enc_out[0], enc_out[1], B, pl.W[6], last_hidden, logits, topk_idx,
                             pool_scores, st);
}

int aasist_forward(aasist_handle* h, const float* x, int32_t B, int32_t L, float* last_hidden, float* logits,
                   int32_t* topk_idx, float* pool_scores, void* workspace, int64_t workspace_bytes, void* stream) {
  return aasist_forward_ex(h, x, B, L, nullptr, last_hidden, logits, topk_idx, pool_scores, workspace,
                           workspace_bytes, stream);
}

int aasist_forward_host(aasist_handle* h, const float* x_host, int32_t B, int32_t L,
                        float* last_hidden_host, float* logits_host, void* stream) {
  if (!h || !x_host || B < 1) {
    set_error("aasist_forward_host: invalid arguments");
    return AASIST_E_INVALID;
  }
  AASIST_ENTER_DEVICE(h);
  int64_t ws_bytes = aasist_workspace_bytes(h, B, L);
  if (ws_bytes < 0) return (int)ws_bytes;
  const int hd = aasist_hidden_dim(h);
  size_t xb = sizeof(float) * (size_t)B * L, ob = sizeof(float) * (size_t)B * (hd + 2);
  size_t dev_need = align256(xb) + align256(ob);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->pin_x_bytes < xb) {
    if (h->pin_x) cudaFreeHost(h->pin_x);
    h->pin_x = nullptr;
    h->pin_x_bytes = 0;
    AASIST_CUDA(cudaMallocHost(&h->pin_x, xb));
    h->pin_x_bytes = xb;
  }
  if (h->pin_out_bytes < ob) {
    if (h->pin_out) cudaFreeHost(h->pin_out);
    h->pin_out = nullptr;
    h->pin_out_bytes = 0;
    AASIST_CUDA(cudaMallocHost(&h->pin_out, ob));
    h->pin_out_bytes = ob;
  }
  if (h->dev_stage_bytes < dev_need) {
    cudaFree(h->dev_stage);
    h->dev_stage = nullptr;
    h->dev_stage_bytes = 0;
    AASIST_CUDA(cudaMalloc(&h->dev_stage, dev_need));
    h->dev_stage_bytes = dev_need;
  }
  // one scratch buffer for every way into the forward (shared with aasist_forward_ex(workspace = NULL))
  if ((rc = ensure_own_ws(h, (size_t)ws_bytes))) return rc;
  char* d = (char*)h->dev_stage;
  float* dx = (float*)d;
  float* dout = (float*)(d + align256(xb));
  // the caller's buffer may be pageable: stage through pinned memory when it is
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, x_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const float* src = x_host;
  if (!pinned) {
    memcpy(h->pin_x, x_host, xb);
    src = h->pin_x;
  }
  if (!h->copy_stream) {
    AASIST_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) AASIST_CUDA(cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
    AASIST_CUDA(cudaEventCreateWithFlags(&h->start_ev, cudaEventDisableTiming));
  }
  // pipeline: a small first piece (its copy is the only one exposed), then the rest of the batch in ONE forward
  // whose copy hides behind the first piece's compute -- large encoder passes are ~3 % more efficient than 128s
  const int first = std::min(B, 128);
  float* d_lh = dout;
  float* d_lg = dout + (size_t)B * hd;
  AASIST_CUDA(cudaEventRecord(h->start_ev, st));                 // staging buffers are free once prior work is done
  AASIST_CUDA(cudaStreamWaitEvent(h->copy_stream, h->start_ev, 0));
  const int piece_b0[2] = {0, first}, piece_nb[2] = {first, B - first};
  for (int c = 0; c < 2; ++c) {                                  // both copies are queued before any compute
    if (piece_nb[c] <= 0) continue;
    AASIST_CUDA(cudaMemcpyAsync(dx + (size_t)piece_b0[c] * L, src + (size_t)piece_b0[c] * L,
                                sizeof(float) * (size_t)piece_nb[c] * L, cudaMemcpyHostToDevice, h->copy_stream));
    AASIST_CUDA(cudaEventRecord(h->copy_done[c], h->copy_stream));
  }
  for (int c = 0; c < 2; ++c) {
    if (piece_nb[c] <= 0) continue;
    const int b0 = piece_b0[c], nb = piece_nb[c];
    AASIST_CUDA(cudaStreamWaitEvent(st, h->copy_done[c], 0));
    if ((rc = aasist_forward(h, dx + (size_t)b0 * L, nb, L, d_lh + (size_t)b0 * hd, d_lg + (size_t)b0 * 2, nullptr,
                             nullptr, h->own_ws, (int64_t)h->own_ws_bytes, stream)))
      return rc;
  }
  AASIST_CUDA(cudaMemcpyAsync(h->pin_out, dout, ob, cudaMemcpyDeviceToHost, st));
  AASIST_CUDA(cudaStreamSynchronize(st));
  if (last_hidden_host) memcpy(last_hidden_host, h->pin_out, sizeof(float) * (size_t)B * hd);
  if (logits_host) memcpy(logits_host, h->pin_out + (size_t)B * hd, sizeof(float) * (size_t)B * 2);
  return AASIST_OK;
}

// ---- the scoring loop: pipelined host -> device -> scores ------------------------------------------
int aasist_score_begin(aasist_handle* h, int64_t capacity, int32_t max_batch, int32_t L, void* stream) {
  if (!h || capacity < 1 || max_batch < 1) {
    set_error("aasist_score_begin: invalid arguments");
    return AASIST_E_INVALID;
  }
  if (!h->finalized) {
    set_error("aasist_score_begin called before aasist_finalize");
    return AASIST_E_STATE;
  }
  AASIST_ENTER_DEVICE(h);
  int64_t ws_bytes = aasist_workspace_bytes(h, max_batch, L);
  if (ws_bytes < 0) return (int)ws_bytes;
  aasist_handle::ScoreStream& sc = h->score;
  if (sc.active) AASIST_CUDA(cudaStreamSynchronize(sc.st));     // an abandoned session: drain it first
  const int hd = aasist_hidden_dim(h);
  const size_t xb = sizeof(float) * (size_t)max_batch * L;
  if (sc.dx_bytes < xb) {
    for (int i = 0; i < 2; ++i) {
      cudaFree(sc.dx[i]);
      sc.dx[i] = nullptr;
    }
    sc.dx_bytes = 0;
    for (int i = 0; i < 2; ++i) AASIST_CUDA(cudaMalloc(&sc.dx[i], xb));
    sc.dx_bytes = xb;
  }
  if (sc.out_cap < (size_t)capacity) {
    cudaFree(sc.d_logits);
    cudaFree(sc.d_hidden);
    sc.d_logits = sc.d_hidden = nullptr;
    sc.out_cap = 0;
    AASIST_CUDA(cudaMalloc(&sc.d_logits, sizeof(float) * 2 * (size_t)capacity));
    AASIST_CUDA(cudaMalloc(&sc.d_hidden, sizeof(float) * hd * (size_t)capacity));
    sc.out_cap = (size_t)capacity;
  }
  for (int i = 0; i < 2; ++i) {
    if (!sc.h2d_done[i]) AASIST_CUDA(cudaEventCreateWithFlags(&sc.h2d_done[i], cudaEventDisableTiming));
    if (!sc.fwd_done[i]) AASIST_CUDA(cudaEventCreateWithFlags(&sc.fwd_done[i], cudaEventDisableTiming));
    sc.used[i] = false;
  }
  if (!h->copy_stream) {
    AASIST_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) AASIST_CUDA(cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
    AASIST_CUDA(cudaEventCreateWithFlags(&h->start_ev, cudaEventDisableTiming));
  }
  if ((rc = ensure_own_ws(h, (size_t)ws_bytes))) return rc;
  sc.st = (cudaStream_t)stream;
  sc.capacity = capacity;
  sc.max_batch = max_batch;
  sc.L = L;
  sc.n = 0;
  sc.slot = 0;
  sc.active = true;
  // the accumulated outputs of a previous session may still be in flight on another stream
  AASIST_CUDA(cudaEventRecord(h->start_ev, sc.st));
  AASIST_CUDA(cudaStreamWaitEvent(h->copy_stream, h->start_ev, 0));
  return AASIST_OK;
}

int aasist_score_submit(aasist_handle* h, const float* x_host, int32_t B) {
  if (!h || !x_host || B < 1) {
    set_error("aasist_score_submit: invalid arguments");
    return AASIST_E_INVALID;
  }
  aasist_handle::ScoreStream& sc = h->score;
  if (!sc.active) {
    set_error("aasist_score_submit without aasist_score_begin");
    return AASIST_E_STATE;
  }
  if (B > sc.max_batch || sc.n + B > sc.capacity) {
    set_error("aasist_score_submit: batch of %d exceeds max_batch %d or the session capacity (%lld of %lld used)", B,
              sc.max_batch, (long long)sc.n, (long long)sc.capacity);
    return AASIST_E_INVALID;
  }
  AASIST_ENTER_DEVICE(h);
  const int slot = sc.slot;
  const int hd = aasist_hidden_dim(h);
  const size_t xb = sizeof(float) * (size_t)B * sc.L;
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, x_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const float* src = x_host;
  if (!pinned) {
    // pageable source: stage through this slot's pinned buffer (wait until its previous H2D has drained)
    if (sc.pin_bytes < sizeof(float) * (size_t)sc.max_batch * sc.L) {
      AASIST_CUDA(cudaStreamSynchronize(h->copy_stream));
      for (int i = 0; i < 2; ++i) {
        if (sc.pin[i]) cudaFreeHost(sc.pin[i]);
        sc.pin[i] = nullptr;
      }
      sc.pin_bytes = 0;
      for (int i = 0; i < 2; ++i) AASIST_CUDA(cudaMallocHost(&sc.pin[i], sizeof(float) * (size_t)sc.max_batch * sc.L));
      sc.pin_bytes = sizeof(float) * (size_t)sc.max_batch * sc.L;
    }
    if (sc.used[slot]) AASIST_CUDA(cudaEventSynchronize(sc.h2d_done[slot]));
    memcpy(sc.pin[slot], x_host, xb);
    src = sc.pin[slot];
  }
  // the device slot is free once the forward that read it (two submits ago) has finished
  if (sc.used[slot]) AASIST_CUDA(cudaStreamWaitEvent(h->copy_stream, sc.fwd_done[slot], 0));
  AASIST_CUDA(cudaMemcpyAsync(sc.dx[slot], src, xb, cudaMemcpyHostToDevice, h->copy_stream));
  AASIST_CUDA(cudaEventRecord(sc.h2d_done[slot], h->copy_stream));
  AASIST_CUDA(cudaStreamWaitEvent(sc.st, sc.h2d_done[slot], 0));
  if ((rc = aasist_forward(h, sc.dx[slot], B, sc.L, sc.d_hidden + (size_t)sc.n * hd, sc.d_logits + (size_t)sc.n * 2,
                           nullptr, nullptr, h->own_ws, (int64_t)h->own_ws_bytes, (void*)sc.st)))
    return rc;
  AASIST_CUDA(cudaEventRecord(sc.fwd_done[slot], sc.st));
  sc.used[slot] = true;
  sc.n += B;
  sc.slot ^= 1;
  return AASIST_OK;
}

int64_t aasist_score_finish(aasist_handle* h, float* logits_out, float* last_hidden_out,
                            const float** logits_dev_out) {
  if (!h) {
    set_error("aasist_score_finish: null handle");
    return AASIST_E_INVALID;
  }
  aasist_handle::ScoreStream& sc = h->score;
  if (!sc.active) {
    set_error("aasist_score_finish without aasist_score_begin");
    return AASIST_E_STATE;
  }
  AASIST_ENTER_DEVICE(h);
  const int hd = aasist_hidden_dim(h);
  if (logits_out && sc.n > 0)      // host or device destination (unified addressing resolves the direction)
    AASIST_CUDA(cudaMemcpyAsync(logits_out, sc.d_logits, sizeof(float) * 2 * (size_t)sc.n, cudaMemcpyDefault, sc.st));
  if (last_hidden_out && sc.n > 0)
    AASIST_CUDA(cudaMemcpyAsync(last_hidden_out, sc.d_hidden, sizeof(float) * hd * (size_t)sc.n, cudaMemcpyDefault,
                                sc.st));
  AASIST_CUDA(cudaStreamSynchronize(sc.st));
  if (logits_dev_out) *logits_dev_out = sc.d_logits;
  sc.active = false;
  return sc.n;
}

// ---- per-stage entry points ---------------------------------------------------------------
int aasist_get_filterbank(aasist_handle* h, float* bank_dev, int32_t* n_filters, int32_t* taps) {
  if (!h || !h->finalized) {
    set_error("aasist_get_filterbank: handle not finalized");
    return AASIST_E_STATE;
  }
  if (n_filters) *n_filters = h->cfg.n_filters;
  if (taps) *taps = h->taps;
  if (bank_dev)
    AASIST_CUDA(cudaMemcpy(bank_dev, h->bank, sizeof(float) * h->cfg.n_filters * h->taps, cudaMemcpyDefault));
  return AASIST_OK;
}

int aasist_frontend(aasist_handle* h, const float* x, int32_t B, int32_t L, float* out, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  if (!h || !h->finalized || !x || !out) {
    set_error("aasist_frontend: invalid arguments or handle not finalized");
    return AASIST_E_STATE;
  }
  AASIST_ENTER_DEVICE(h);
  Plan pl;
  if ((rc = make_plan(h, L, pl))) return rc;
  (void)workspace;
  (void)workspace_bytes;
  if (h->cfg.kind == AASIST_KIND_ROBUST) return launch_frontend_strided_f32(h, x, B, L, out, 0, 0, (cudaStream_t)stream);
  if (h->cfg.precision != AASIST_PREC_FP32) return tc_frontend_to_f32(h, x, B, L, out, 0, 0, (cudaStream_t)stream);
  return launch_frontend_f32(h, x, B, L, out, 0, 0, (cudaStream_t)stream);
}

int aasist_encoder_block(aasist_handle* h, int32_t enc, int32_t index, const float* in, int32_t B, int32_t W,
                         float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !h->finalized || !in || !out || enc < 0 || enc >= h->n_encoders || index < 0 || index > 5) {
    set_error("aasist_encoder_block: invalid arguments or handle not finalized");
    return AASIST_E_STATE;
  }
  AASIST_ENTER_DEVICE(h);
  if (tc_encoder(h))
    return tc_block_f32io(h, enc, index, in, B, W, out, workspace, workspace_bytes, (cudaStream_t)stream);
  const int co = h->cfg.enc_channels[index][1];
  size_t need = sizeof(float) * (h->cfg.encoder == AASIST_ENC_RES2NET ? res2_block_scratch_floats(h->res2[index], B, W)
                                                                      : (size_t)B * co * 24 * W);
  if (!workspace || (size_t)workspace_bytes < need) {
    set_error("aasist_encoder_block: workspace needs %zu bytes", need);
    return AASIST_E_WORKSPACE;
  }
  if (h->cfg.encoder == AASIST_ENC_RES2NET)
    return launch_res2_block(h, h->res2[index], in, B, W, (float*)workspace, out, (cudaStream_t)stream);
  if (h->cfg.encoder == AASIST_ENC_RESIDUAL33)
    return launch_block33_f32(h, h->blocks33[index], in, B, W, (float*)workspace, out, (cudaStream_t)stream);
  return launch_block_f32(h, h->blocks[enc][index], in, B, W, (float*)workspace, out, (cudaStream_t)stream);
}

int aasist_graph(aasist_handle* h, const float* e, const float* e2, int32_t B, int32_t NT, float* last_hidden,
                 float* logits, int32_t* topk_idx, float* pool_scores, void* stream) {
  if (!h || !h->finalized || !e || !last_hidden || !logits || B < 1 || NT < 1) {
    set_error("aasist_graph: invalid arguments or handle not finalized");
    return AASIST_E_STATE;
  }
  AASIST_ENTER_DEVICE(h);
  if (h->cfg.kind != AASIST_KIND_RAWGAT_ST)
    return launch_graph_aasist(h, e, B, NT, nullptr, last_hidden, logits, topk_idx, pool_scores, (cudaStream_t)stream);
  if (!e2) {
    set_error("aasist_graph: RawGAT-ST needs both encoder outputs");
    return AASIST_E_INVALID;
  }
  return launch_graph_rawgat(h, e, e2, B, NT, last_hidden, logits, topk_idx, pool_scores, (cudaStream_t)stream);
}

int64_t aasist_launch_count(const aasist_handle* h) { return h ? h->launches : 0; }

int aasist_input_range_exceeded(aasist_handle* h, int32_t reset) {
  if (!h || !h->range_flag) return 0;
  const int v = *reinterpret_cast<volatile int*>(h->range_flag);
  if (reset) *h->range_flag = 0;
  return v != 0;
}

int aasist_profile_enable(aasist_handle* h, int32_t enable) {
  if (!h) return AASIST_E_INVALID;
  h->profiling = enable != 0;
  return AASIST_OK;
}

int aasist_profile_report(aasist_handle* h, char* buf, int64_t buf_bytes, int32_t reset) {
  if (!h || !buf || buf_bytes < 8) {
    set_error("aasist_profile_report: invalid arguments");
    return AASIST_E_INVALID;
  }
  AASIST_ENTER_DEVICE(h);
  AASIST_CUDA(cudaDeviceSynchronize());
  for (auto& sp : h->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      auto& t = h->prof_totals[sp.name];
      t.first += 1;
      t.second += ms;
    } else {
      cudaGetLastError();
    }
    h->prof_pool.push_back(sp.a);
    h->prof_pool.push_back(sp.b);
  }
  h->prof_pending.clear();
  std::string js = "[";
  bool first = true;
  for (auto& kv : h->prof_totals) {
    char item[256];
    snprintf(item, sizeof(item), "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ",
             kv.first.c_str(), (long long)kv.second.first, kv.second.second);
    js += item;
    first = false;
  }
  js += "]";
  if ((int64_t)js.size() + 1 > buf_bytes) {
    set_error("aasist_profile_report: buffer too small (%zu bytes needed)", js.size() + 1);
    return AASIST_E_INVALID;
  }
  memcpy(buf, js.c_str(), js.size() + 1);
  if (reset) h->prof_totals.clear();
  return (int)js.size();
}

}  // extern "C"
#pragma GCC visibility pop
