// onnx_loader.cu -- reads the model file the reference actually loads.
//
// OrtKoko::new(model_path) hands the downloaded ONNX file to ONNX Runtime
// (/root/reference/kokorox/src/onn/ort_base.rs:27-33 `commit_from_file(model_path)`; file list
// kokorox/src/utils/hf_cache.rs:135-144: model.onnx, model_fp16, model_quantized / uint8 / q8f16, q4 / q4f16;
// default URL kokorox/src/tts/koko.rs:57).  kkx_create accepts the same file: this unit parses the protobuf wire
// format (no protobuf / onnx dependency; field numbers from onnx.proto3), recovers the Kokoro-82M state dict and
// hands it to WeightSet::load as named fp32 tensors.  Host-only code (the .cu suffix just puts it in the library).
//
// What a torch.onnx export of Kokoro looks like, and how each kind of tensor is found:
//   * embeddings, LayerNorm / channel-LayerNorm affine, biases, Snake alphas and un-weight-normed convs keep their
//     state-dict names, possibly behind a wrapper prefix ("kmodel.", "model.", "module.") -> taken BY NAME;
//   * nn.Linear weights become anonymous, transposed MatMul operands ("onnx::MatMul_1234", [K, N]); weight-normed
//     conv weights are constant-folded into anonymous "onnx::Conv_1234"; nn.LSTM parameters become the W / R / B
//     operands of an LSTM node in ONNX gate order i,o,f,c -> named THROUGH THE GRAPH: the consuming node's own name
//     carries the module path ("/bert/encoder/albert_layer_groups.0/albert_layers.0/attention/query/MatMul_3" ->
//     "bert.encoder.albert_layer_groups.0.albert_layers.0.attention.query"), and where it does not, the module is
//     taken from the bias initialiser the node (Conv) or the Add behind it (MatMul) consumes;
//   * weight-norm left unfolded ("...weight_g"/"...weight_v" or "...parametrizations.weight.original0/1") is folded
//     here: w = g * v / ||v|| over every dim but 0;
//   * fp16 / bf16 / double payloads are widened; dynamic-quantisation graphs (MatMulInteger, ConvInteger,
//     DynamicQuantizeLSTM, DequantizeLinear on initialisers) are dequantised with their scale / zero-point
//     initialisers; 4-bit block-quantised MatMulNBits weights are unpacked.
// Anything still missing afterwards is an IoError that lists the missing tensors and the anonymous initialisers
// that were left over -- nothing is guessed by shape.
#include "model.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <sstream>

namespace kkx {

// ------------------------------------------------------------------------------------------------ tensor specs
// (name, shape) of every tensor WeightSet::load reads: SURVEY.md A.12 / A.13, mirrors kokorox_b200/weightfile.py
// weight_specs() (tests/test_onnx_loader.py compares the two lists through kkx_test_tensor_specs).
std::vector<std::pair<std::string, std::vector<int>>> kokoro_tensor_specs() {
  std::vector<std::pair<std::string, std::vector<int>>> s;
  auto add = [&](const std::string& n, std::vector<int> shp) { s.push_back({n, std::move(shp)}); };
  auto lstm = [&](const std::string& p, int in) {
    for (const char* sfx : {"", "_reverse"}) {
      add(p + ".weight_ih_l0" + sfx, {1024, in}); add(p + ".weight_hh_l0" + sfx, {1024, 256});
      add(p + ".bias_ih_l0" + sfx, {1024}); add(p + ".bias_hh_l0" + sfx, {1024});
    }
  };
  auto blk = [&](const std::string& p, int ci, int co, bool up) {
    add(p + ".conv1.weight", {co, ci, 3}); add(p + ".conv1.bias", {co});
    add(p + ".conv2.weight", {co, co, 3}); add(p + ".conv2.bias", {co});
    add(p + ".norm1.fc.weight", {2 * ci, 128}); add(p + ".norm1.fc.bias", {2 * ci});
    add(p + ".norm2.fc.weight", {2 * co, 128}); add(p + ".norm2.fc.bias", {2 * co});
    if (ci != co) add(p + ".conv1x1.weight", {co, ci, 1});
    if (up) { add(p + ".pool.weight", {ci, 1, 3}); add(p + ".pool.bias", {ci}); }
  };
  auto arb = [&](const std::string& p, int c, int k) {
    for (int j = 0; j < 3; j++) {
      const std::string sj = std::to_string(j);
      add(p + ".convs1." + sj + ".weight", {c, c, k}); add(p + ".convs1." + sj + ".bias", {c});
      add(p + ".convs2." + sj + ".weight", {c, c, k}); add(p + ".convs2." + sj + ".bias", {c});
      add(p + ".adain1." + sj + ".fc.weight", {2 * c, 128}); add(p + ".adain1." + sj + ".fc.bias", {2 * c});
      add(p + ".adain2." + sj + ".fc.weight", {2 * c, 128}); add(p + ".adain2." + sj + ".fc.bias", {2 * c});
      add(p + ".alpha1." + sj, {1, c, 1}); add(p + ".alpha2." + sj, {1, c, 1});
    }
  };
  const std::string E = "bert.embeddings.", L = "bert.encoder.albert_layer_groups.0.albert_layers.0.";
  add(E + "word_embeddings.weight", {178, 128}); add(E + "position_embeddings.weight", {512, 128});
  add(E + "token_type_embeddings.weight", {2, 128});
  add(E + "LayerNorm.weight", {128}); add(E + "LayerNorm.bias", {128});
  add("bert.encoder.embedding_hidden_mapping_in.weight", {768, 128});
  add("bert.encoder.embedding_hidden_mapping_in.bias", {768});
  for (const char* nm : {"query", "key", "value", "dense"}) {
    add(L + "attention." + nm + ".weight", {768, 768}); add(L + "attention." + nm + ".bias", {768});
  }
  add(L + "attention.LayerNorm.weight", {768}); add(L + "attention.LayerNorm.bias", {768});
  add(L + "ffn.weight", {2048, 768}); add(L + "ffn.bias", {2048});
  add(L + "ffn_output.weight", {768, 2048}); add(L + "ffn_output.bias", {768});
  add(L + "full_layer_layer_norm.weight", {768}); add(L + "full_layer_layer_norm.bias", {768});
  add("bert_encoder.weight", {512, 768}); add("bert_encoder.bias", {512});
  add("text_encoder.embedding.weight", {178, 512});
  for (int i = 0; i < 3; i++) {
    const std::string p = "text_encoder.cnn." + std::to_string(i);
    add(p + ".0.weight", {512, 512, 5}); add(p + ".0.bias", {512});
    add(p + ".1.gamma", {512}); add(p + ".1.beta", {512});
  }
  lstm("text_encoder.lstm", 512);
  for (int i = 0; i < 3; i++) {
    lstm("predictor.text_encoder.lstms." + std::to_string(2 * i), 640);
    add("predictor.text_encoder.lstms." + std::to_string(2 * i + 1) + ".fc.weight", {1024, 128});
    add("predictor.text_encoder.lstms." + std::to_string(2 * i + 1) + ".fc.bias", {1024});
  }
  lstm("predictor.lstm", 640);
  add("predictor.duration_proj.linear_layer.weight", {50, 512});
  add("predictor.duration_proj.linear_layer.bias", {50});
  lstm("predictor.shared", 640);
  for (const char* br : {"F0", "N"}) {
    const std::string p = std::string("predictor.") + br;
    blk(p + ".0", 512, 512, false); blk(p + ".1", 512, 256, true); blk(p + ".2", 256, 256, false);
  }
  add("predictor.F0_proj.weight", {1, 256, 1}); add("predictor.F0_proj.bias", {1});
  add("predictor.N_proj.weight", {1, 256, 1}); add("predictor.N_proj.bias", {1});
  blk("decoder.encode", 514, 1024, false);
  for (int i = 0; i < 3; i++) blk("decoder.decode." + std::to_string(i), 1090, 1024, false);
  blk("decoder.decode.3", 1090, 512, true);
  add("decoder.F0_conv.weight", {1, 1, 3}); add("decoder.F0_conv.bias", {1});
  add("decoder.N_conv.weight", {1, 1, 3}); add("decoder.N_conv.bias", {1});
  add("decoder.asr_res.0.weight", {64, 512, 1}); add("decoder.asr_res.0.bias", {64});
  const std::string G = "decoder.generator.";
  add(G + "m_source.l_linear.weight", {1, 9}); add(G + "m_source.l_linear.bias", {1});
  add(G + "noise_convs.0.weight", {256, 22, 12}); add(G + "noise_convs.0.bias", {256});
  add(G + "noise_convs.1.weight", {128, 22, 1}); add(G + "noise_convs.1.bias", {128});
  arb(G + "noise_res.0", 256, 7); arb(G + "noise_res.1", 128, 11);
  add(G + "ups.0.weight", {512, 256, 20}); add(G + "ups.0.bias", {256});
  add(G + "ups.1.weight", {256, 128, 12}); add(G + "ups.1.bias", {128});
  const int rk[3] = {3, 7, 11};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) arb(G + "resblocks." + std::to_string(i * 3 + j), i == 0 ? 256 : 128, rk[j]);
  add(G + "conv_post.weight", {22, 128, 7}); add(G + "conv_post.bias", {22});
  return s;
}

namespace {

// ------------------------------------------------------------------------------------------------ protobuf wire
struct Span { const uint8_t* p = nullptr; size_t n = 0; };
struct Field { uint32_t no = 0, wt = 0; uint64_t v = 0; Span s; };

class Pb {
 public:
  Pb(const uint8_t* p, size_t n) : p_(p), e_(p + n) {}
  explicit Pb(Span s) : p_(s.p), e_(s.p + s.n) {}
  bool next(Field& f) {
    if (p_ >= e_) return false;
    const uint64_t key = varint();
    f.no = (uint32_t)(key >> 3); f.wt = (uint32_t)(key & 7); f.v = 0; f.s = Span();
    switch (f.wt) {
      case 0: f.v = varint(); break;
      case 1: need(8); memcpy(&f.v, p_, 8); p_ += 8; break;
      case 2: { const uint64_t ln = varint(); need(ln); f.s.p = p_; f.s.n = (size_t)ln; p_ += ln; break; }
      case 5: { need(4); uint32_t x; memcpy(&x, p_, 4); f.v = x; p_ += 4; break; }
      default: throw IoError("ONNX: unsupported protobuf wire type " + std::to_string(f.wt));
    }
    return true;
  }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      need(1);
      const uint8_t b = *p_++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (b < 0x80) return v;
    }
    throw IoError("ONNX: malformed varint");
  }
  bool done() const { return p_ >= e_; }

 private:
  void need(uint64_t k) { if (k > (uint64_t)(e_ - p_)) throw IoError("ONNX: truncated protobuf field"); }
  const uint8_t* p_; const uint8_t* e_;
};

std::string str(Span s) { return std::string(reinterpret_cast<const char*>(s.p), s.n); }

// ------------------------------------------------------------------------------------------------ ONNX objects
enum { DT_F32 = 1, DT_U8 = 2, DT_I8 = 3, DT_U16 = 4, DT_I16 = 5, DT_I32 = 6, DT_I64 = 7, DT_BOOL = 9, DT_F16 = 10,
       DT_F64 = 11, DT_U32 = 12, DT_U64 = 13, DT_BF16 = 16 };

struct OTensor {
  std::string name;
  std::vector<long long> dims;
  int dtype = DT_F32;
  Span raw;                          // raw_data (little endian)
  std::vector<float> f32;            // float_data
  std::vector<long long> ints;       // int32_data / int64_data (sign-extended)
  std::vector<double> f64;           // double_data
  bool external = false;
  size_t numel() const { size_t n = 1; for (long long d : dims) n *= (size_t)std::max<long long>(d, 0); return n; }
};

struct ONode {
  std::string op, name;
  std::vector<std::string> in, out;
  std::map<std::string, long long> ints;
  std::map<std::string, std::vector<long long>> int_lists;
  std::map<std::string, OTensor> tensors;
};

void packed_varints(const Field& f, std::vector<long long>& out) {
  if (f.wt == 0) { out.push_back((long long)f.v); return; }
  Pb p(f.s);
  while (!p.done()) out.push_back((long long)p.varint());
}

OTensor parse_tensor(Span s) {
  OTensor t;
  Pb p(s);
  Field f;
  while (p.next(f)) {
    switch (f.no) {
      case 1: packed_varints(f, t.dims); break;
      case 2: t.dtype = (int)f.v; break;
      case 4:
        if (f.wt == 2) { const size_t k = f.s.n / 4, o = t.f32.size(); t.f32.resize(o + k); memcpy(t.f32.data() + o, f.s.p, k * 4); }
        else { float x; const uint32_t u = (uint32_t)f.v; memcpy(&x, &u, 4); t.f32.push_back(x); }
        break;
      case 5: { std::vector<long long> v; packed_varints(f, v); for (long long x : v) t.ints.push_back((long long)(int32_t)(uint32_t)x); break; }
      case 7: packed_varints(f, t.ints); break;
      case 8: t.name = str(f.s); break;
      case 9: t.raw = f.s; break;
      case 10:
        if (f.wt == 2) { const size_t k = f.s.n / 8, o = t.f64.size(); t.f64.resize(o + k); memcpy(t.f64.data() + o, f.s.p, k * 8); }
        else { double x; memcpy(&x, &f.v, 8); t.f64.push_back(x); }
        break;
      case 14: if (f.v == 1) t.external = true; break;
      default: break;
    }
  }
  for (long long d : t.dims)
    if (d < 0 || d > (1LL << 32)) throw IoError("ONNX: tensor " + t.name + " has a bad dimension");
  return t;
}

ONode parse_node(Span s) {
  ONode n;
  Pb p(s);
  Field f;
  while (p.next(f)) {
    if (f.no == 1 && f.wt == 2) n.in.push_back(str(f.s));
    else if (f.no == 2 && f.wt == 2) n.out.push_back(str(f.s));
    else if (f.no == 3 && f.wt == 2) n.name = str(f.s);
    else if (f.no == 4 && f.wt == 2) n.op = str(f.s);
    else if (f.no == 5 && f.wt == 2) {          // AttributeProto: name 1, i 3, t 5, ints 8
      Pb a(f.s);
      Field g;
      std::string an; bool has_i = false; long long iv = 0; Span tv; std::vector<long long> il;
      while (a.next(g)) {
        if (g.no == 1 && g.wt == 2) an = str(g.s);
        else if (g.no == 3 && g.wt == 0) { has_i = true; iv = (long long)g.v; }
        else if (g.no == 5 && g.wt == 2) tv = g.s;
        else if (g.no == 8) packed_varints(g, il);
      }
      if (has_i) n.ints[an] = iv;
      if (!il.empty()) n.int_lists[an] = il;
      if (tv.p) n.tensors[an] = parse_tensor(tv);
    }
  }
  return n;
}

float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000) << 16;
  uint32_t e = (h >> 10) & 0x1F, m = h & 0x3FF, u;
  if (e == 0) {
    if (m == 0) u = sign;
    else {                                 // subnormal: renormalise
      int sh = 0;
      while (!(m & 0x400)) { m <<= 1; sh++; }
      m &= 0x3FF;
      u = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
    }
  } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
  else u = sign | ((e - 15 + 127) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// A dense host tensor in fp32 (integers are converted value by value: quantised payloads keep their codes).
struct FT { std::vector<long long> dims; std::vector<float> v; bool ok = false; };

FT to_float(const OTensor& t) {
  if (t.external) throw IoError("ONNX: tensor " + t.name + " uses external data (a second file), which is not supported");
  FT o;
  o.dims = t.dims;
  const size_t n = t.numel();
  o.v.resize(n);
  auto raw_as = [&](size_t elt) { if (t.raw.n != n * elt) throw IoError("ONNX: tensor " + t.name + ": payload size does not match its shape"); };
  const uint8_t* r = t.raw.p;
  switch (t.dtype) {
    case DT_F32:
      if (r) { raw_as(4); memcpy(o.v.data(), r, n * 4); }
      else if (t.f32.size() == n) o.v = t.f32;
      else throw IoError("ONNX: tensor " + t.name + ": float_data size does not match its shape");
      break;
    case DT_F16: case DT_BF16:
      for (size_t i = 0; i < n; i++) {
        uint16_t h;
        if (r) { if (i == 0) raw_as(2); memcpy(&h, r + 2 * i, 2); }
        else if (t.ints.size() == n) h = (uint16_t)t.ints[i];
        else throw IoError("ONNX: tensor " + t.name + ": half payload missing");
        if (t.dtype == DT_F16) o.v[i] = half_to_float(h);
        else { const uint32_t u = (uint32_t)h << 16; memcpy(&o.v[i], &u, 4); }
      }
      break;
    case DT_F64:
      for (size_t i = 0; i < n; i++) {
        double d;
        if (r) { if (i == 0) raw_as(8); memcpy(&d, r + 8 * i, 8); }
        else if (t.f64.size() == n) d = t.f64[i];
        else throw IoError("ONNX: tensor " + t.name + ": double payload missing");
        o.v[i] = (float)d;
      }
      break;
    case DT_U8: case DT_I8: case DT_BOOL:
      for (size_t i = 0; i < n; i++) {
        if (r) { if (i == 0) raw_as(1); o.v[i] = t.dtype == DT_I8 ? (float)(int8_t)r[i] : (float)r[i]; }
        else if (t.ints.size() == n) o.v[i] = (float)t.ints[i];
        else throw IoError("ONNX: tensor " + t.name + ": int8 payload missing");
      }
      break;
    case DT_I32: case DT_I64: case DT_I16: case DT_U16: case DT_U32: case DT_U64: {
      const size_t elt = (t.dtype == DT_I64 || t.dtype == DT_U64) ? 8 : (t.dtype == DT_I32 || t.dtype == DT_U32) ? 4 : 2;
      for (size_t i = 0; i < n; i++) {
        if (r) {
          if (i == 0) raw_as(elt);
          if (elt == 8) { int64_t x; memcpy(&x, r + 8 * i, 8); o.v[i] = (float)x; }
          else if (elt == 4) { int32_t x; memcpy(&x, r + 4 * i, 4); o.v[i] = t.dtype == DT_U32 ? (float)(uint32_t)x : (float)x; }
          else { int16_t x; memcpy(&x, r + 2 * i, 2); o.v[i] = t.dtype == DT_U16 ? (float)(uint16_t)x : (float)x; }
        } else if (t.ints.size() == n) o.v[i] = (float)t.ints[i];
        else throw IoError("ONNX: tensor " + t.name + ": integer payload missing");
      }
      break;
    }
    default: throw IoError("ONNX: tensor " + t.name + ": unsupported data_type " + std::to_string(t.dtype));
  }
  o.ok = true;
  return o;
}

bool is_float_type(int dt) { return dt == DT_F32 || dt == DT_F16 || dt == DT_BF16 || dt == DT_F64; }
bool is_quant_type(int dt) { return dt == DT_U8 || dt == DT_I8; }

// ------------------------------------------------------------------------------------------------ the graph
struct Graph {
  std::map<std::string, OTensor> init;        // initialisers + Constant outputs, by value name
  std::vector<ONode> nodes;
  std::map<std::string, int> producer;        // value name -> node
  std::multimap<std::string, int> consumers;  // value name -> nodes
};

Graph parse_model(const std::vector<char>& bytes, const std::string& path) {
  Pb m(reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size());
  Field f;
  Span graph;
  try {
    while (m.next(f))
      if (f.no == 7 && f.wt == 2) graph = f.s;
  } catch (const IoError&) {
    throw IoError(path + ": neither a KKXW0001 weight file nor an ONNX model (protobuf does not parse)");
  }
  if (!graph.p) throw IoError(path + ": neither a KKXW0001 weight file nor an ONNX model (no GraphProto)");
  Graph g;
  Pb gp(graph);
  while (gp.next(f)) {
    if (f.no == 5 && f.wt == 2) { OTensor t = parse_tensor(f.s); const std::string nm = t.name; g.init[nm] = std::move(t); }
    else if (f.no == 1 && f.wt == 2) g.nodes.push_back(parse_node(f.s));
  }
  for (int i = 0; i < (int)g.nodes.size(); i++) {
    ONode& n = g.nodes[i];
    if (n.op == "Constant" && !n.out.empty()) {
      auto it = n.tensors.find("value");
      if (it != n.tensors.end()) { OTensor t = it->second; t.name = n.out[0]; g.init[n.out[0]] = std::move(t); }
    }
    for (auto& o : n.out) g.producer[o] = i;
    for (auto& in : n.in) if (!in.empty()) g.consumers.insert({in, i});
  }
  return g;
}

// ------------------------------------------------------------------------------------------------ naming
const char* kTop[] = {"bert_encoder", "bert", "predictor", "text_encoder", "decoder"};

// "kmodel.bert.embeddings.x" / "module.decoder.y" -> "bert.embeddings.x" / "decoder.y"; "" when the name is not
// under one of Kokoro's five top-level groups.
std::string canonical(const std::string& name) {
  size_t best = std::string::npos;
  for (const char* t : kTop) {
    const std::string key = t;
    size_t pos = 0;
    while ((pos = name.find(key, pos)) != std::string::npos) {
      const bool left = pos == 0 || name[pos - 1] == '.';
      const bool right = pos + key.size() < name.size() && name[pos + key.size()] == '.';
      if (left && right) { best = std::min(best, pos); break; }
      pos += key.size();
    }
  }
  if (best == std::string::npos) return "";
  std::string c = name.substr(best);
  // DataParallel leftovers inside the path ("decoder.module.generator...")
  size_t q;
  while ((q = c.find(".module.")) != std::string::npos) c.erase(q, 7);
  return c;
}

// torch.onnx scope path -> module path.  "/text_encoder/cnn.0/cnn.0.0/Conv" -> "text_encoder.cnn.0.0": every scope
// level is the module's name with the parents' leading atoms removed but numeric atoms kept together with the first
// non-numeric atom before them (torch.onnx _unqualified_variable_name), so a level that extends the previous one
// ("cnn.0" then "cnn.0.0") replaces it.  The last segment is the op ("Conv", "MatMul_3", "MatMul_quant").
std::string module_from_node_name(const std::string& node_name) {
  if (node_name.empty() || node_name[0] != '/') return "";
  std::vector<std::string> seg;
  size_t i = 1;
  while (i <= node_name.size()) {
    const size_t j = std::min(node_name.find('/', i), node_name.size());
    seg.push_back(node_name.substr(i, j - i));
    i = j + 1;
  }
  if (seg.size() < 2) return "";
  seg.pop_back();
  std::vector<std::string> kept;
  for (auto& s : seg) {
    if (s.empty()) continue;
    if (!kept.empty() && s.size() > kept.back().size() && s.compare(0, kept.back().size(), kept.back()) == 0 &&
        s[kept.back().size()] == '.')
      kept.back() = s;
    else kept.push_back(s);
  }
  std::string path;
  for (auto& s : kept) path += (path.empty() ? "" : ".") + s;
  return canonical(path);
}

std::string strip_suffix(const std::string& s, const std::string& suf) {
  return s.size() > suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0 ? s.substr(0, s.size() - suf.size()) : "";
}

// ------------------------------------------------------------------------------------------------ the reader
class Reader {
 public:
  Reader(Graph&& g, WeightFile& out) : g_(std::move(g)), out_(out) {
    for (auto& sp : kokoro_tensor_specs()) { need_[sp.first] = sp.second; order_.push_back(sp.first); }
  }

  void run(const std::string& path) {
    by_name();
    by_graph();
    fold_weight_norm();
    std::vector<std::string> missing;
    for (auto& n : order_)
      if (!out_.has(n)) missing.push_back(n);
    if (!missing.empty()) {
      std::ostringstream os;
      os << path << ": ONNX model does not resolve to Kokoro-82M: " << missing.size() << " tensors missing (";
      for (size_t i = 0; i < missing.size() && i < 64; i++) os << (i ? ", " : "") << missing[i];
      if (missing.size() > 64) os << ", ...";
      os << ")";
      std::vector<std::string> left;
      for (auto& kv : g_.init)
        if (!used_.count(kv.first) && kv.second.numel() >= 64 && (is_float_type(kv.second.dtype) || is_quant_type(kv.second.dtype)) &&
            canonical(kv.first).empty())      // named tensors the backend does not use (bert.pooler) are not leftovers
          left.push_back(kv.first);
      if (!left.empty()) {
        os << "; " << left.size() << " weight-sized initialisers were not placed (";
        for (size_t i = 0; i < left.size() && i < 32; i++) os << (i ? ", " : "") << left[i];
        if (left.size() > 32) os << ", ...";
        os << ")";
      }
      if (!notes_.empty()) os << "; " << notes_;
      throw IoError(os.str());
    }
    // drop the un-folded weight-norm halves and anything else that is not part of the spec
    for (auto& n : out_.names())
      if (!need_.count(n)) out_.erase(n);
  }

 private:
  // ---- constant values reachable from initialisers through shape-preserving / dequantising nodes
  bool is_const(const std::string& v) const { return g_.init.count(v) != 0; }

  FT value(const std::string& v, int depth = 0) {
    FT none;
    if (v.empty() || depth > 6) return none;
    auto it = g_.init.find(v);
    if (it != g_.init.end()) {
      // a quantised initialiser with "<base>_scale" / "<base>_zero_point" siblings (ORT quantiser naming)
      if (is_quant_type(it->second.dtype)) {
        const std::string base = strip_suffix(v, "_quantized");
        if (!base.empty() && g_.init.count(base + "_scale")) {
          FT q = to_float(it->second);
          FT sc = to_float(g_.init.at(base + "_scale"));
          FT zp;
          if (g_.init.count(base + "_zero_point")) zp = to_float(g_.init.at(base + "_zero_point"));
          used_.insert(v); used_.insert(base + "_scale"); used_.insert(base + "_zero_point");
          return dequant(q, sc, zp, sc.v.size() > 1 ? guess_axis(q, sc) : 0);
        }
      }
      used_.insert(v);
      return to_float(it->second);
    }
    auto pit = g_.producer.find(v);
    if (pit == g_.producer.end()) return none;
    const ONode& n = g_.nodes[pit->second];
    if (n.op == "DequantizeLinear" && n.in.size() >= 2) {
      FT q = value(n.in[0], depth + 1), sc = value(n.in[1], depth + 1), zp;
      if (n.in.size() > 2 && !n.in[2].empty()) zp = value(n.in[2], depth + 1);
      if (!q.ok || !sc.ok) return none;
      long long axis = n.ints.count("axis") ? n.ints.at("axis") : 1;
      if (axis < 0) axis += (long long)q.dims.size();
      return dequant(q, sc, zp, (int)axis);
    }
    if (n.op == "Cast" || n.op == "Identity") return n.in.empty() ? none : value(n.in[0], depth + 1);
    if (n.op == "Transpose" && !n.in.empty()) {
      FT x = value(n.in[0], depth + 1);
      if (!x.ok || x.dims.size() != 2) return none;
      auto pi = n.int_lists.find("perm");
      if (pi != n.int_lists.end() && pi->second.size() == 2 && pi->second[0] == 0) return x;   // identity perm
      return transpose2(x);
    }
    return none;
  }

  static FT transpose2(const FT& x) {
    FT o;
    o.ok = true;
    const size_t R = (size_t)x.dims[0], C = (size_t)x.dims[1];
    o.dims = {x.dims[1], x.dims[0]};
    o.v.resize(R * C);
    for (size_t r = 0; r < R; r++)
      for (size_t c = 0; c < C; c++) o.v[c * R + r] = x.v[r * C + c];
    return o;
  }

  static int guess_axis(const FT& q, const FT& sc) {
    for (int a = 0; a < (int)q.dims.size(); a++)
      if ((size_t)q.dims[a] == sc.v.size()) return a;
    return 0;
  }

  // (q - zp) * scale; scale / zp scalar or per-slice along `axis`
  static FT dequant(const FT& q, const FT& sc, const FT& zp, int axis) {
    FT o;
    o.ok = true; o.dims = q.dims; o.v.resize(q.v.size());
    size_t inner = 1;
    for (size_t d = (size_t)axis + 1; d < q.dims.size(); d++) inner *= (size_t)q.dims[d];
    const size_t na = sc.v.size() > 1 && axis < (int)q.dims.size() ? (size_t)q.dims[axis] : 1;
    if (sc.v.size() > 1 && sc.v.size() != na) throw IoError("ONNX: per-channel scale does not match the quantised tensor");
    for (size_t i = 0; i < q.v.size(); i++) {
      const size_t a = sc.v.size() > 1 ? (i / inner) % na : 0;
      const float z = zp.ok ? zp.v[zp.v.size() > 1 ? a : 0] : 0.f;
      o.v[i] = (q.v[i] - z) * sc.v[a];
    }
    return o;
  }

  // ---- placement
  bool wanted(const std::string& name) const { return need_.count(name) != 0; }

  // store `t` under canonical `name` if the spec wants it and the element count matches (shape as the spec says:
  // exporters store e.g. Snake alphas as [C] or [1,C,1]); first writer wins unless `force`
  bool place(const std::string& name, FT&& t, bool force = false) {
    auto it = need_.find(name);
    if (it == need_.end()) return false;
    size_t want = 1;
    for (int d : it->second) want *= (size_t)d;
    if (t.v.size() != want) {
      notes_ += (notes_.empty() ? "" : "; ") + name + ": found " + std::to_string(t.v.size()) + " elements, expected " + std::to_string(want);
      return false;
    }
    if (out_.has(name) && !force) return true;
    out_.put(name, it->second, std::move(t.v));
    return true;
  }

  void by_name() {
    for (auto& kv : g_.init) {
      const OTensor& t = kv.second;
      if (!is_float_type(t.dtype)) continue;
      const std::string c = canonical(kv.first);
      if (c.empty()) continue;
      if (wanted(c)) {
        FT f = to_float(t);
        if (place(c, std::move(f))) used_.insert(kv.first);
      } else if (strip_suffix(c, ".weight_g").size() || strip_suffix(c, ".weight_v").size() ||
                 c.find(".parametrizations.weight.original") != std::string::npos) {
        FT f = to_float(t);
        std::vector<int> shp;
        for (long long d : f.dims) shp.push_back((int)d);
        out_.put(c, shp, std::move(f.v));     // kept until fold_weight_norm
        used_.insert(kv.first);
      }
    }
    // named quantised weights ("decoder...weight_quantized" + "_scale" + "_zero_point")
    for (auto& kv : g_.init) {
      if (!is_quant_type(kv.second.dtype)) continue;
      const std::string base = strip_suffix(kv.first, "_quantized");
      if (base.empty()) continue;
      const std::string c = canonical(base);
      if (c.empty() || !wanted(c) || out_.has(c)) continue;
      FT f = value(kv.first);
      if (f.ok) place(c, std::move(f));
    }
  }

  // module of a compute node: its own scope path, else the bias it (or the Add behind it) consumes
  std::string module_of(const ONode& n) {
    std::string m = module_from_node_name(n.name);
    if (!m.empty()) return m;
    auto from_bias = [&](const std::string& v) -> std::string {
      if (!is_const(v)) return "";
      const std::string c = canonical(v);
      return strip_suffix(c, ".bias");
    };
    if ((n.op == "Conv" || n.op == "ConvTranspose" || n.op == "Gemm") && n.in.size() > 2) {
      m = from_bias(n.in[2]);
      if (!m.empty()) return m;
    }
    if (!n.out.empty()) {     // MatMul -> [Cast ->] [Mul ->] Add(bias)
      std::string v = n.out[0];
      for (int hop = 0; hop < 4 && !v.empty(); hop++) {
        auto range = g_.consumers.equal_range(v);
        std::string next;
        for (auto it = range.first; it != range.second; ++it) {
          const ONode& c = g_.nodes[it->second];
          if (c.op == "Add")
            for (auto& in : c.in) { m = from_bias(in); if (!m.empty()) return m; }
          if ((c.op == "Cast" || c.op == "Mul" || c.op == "Reshape") && !c.out.empty()) next = c.out[0];
        }
        v = next;
      }
    }
    return "";
  }

  // bias of a Linear exported as MatMul + Add: the Add's constant operand (named or anonymous)
  void place_linear_bias(const ONode& n, const std::string& mod) {
    if (out_.has(mod + ".bias") || !wanted(mod + ".bias") || n.out.empty()) return;
    std::string v = n.out[0];
    for (int hop = 0; hop < 4 && !v.empty(); hop++) {
      auto range = g_.consumers.equal_range(v);
      std::string next;
      for (auto it = range.first; it != range.second; ++it) {
        const ONode& c = g_.nodes[it->second];
        if (c.op == "Add")
          for (auto& in : c.in)
            if (in != v) { FT b = value(in); if (b.ok && place(mod + ".bias", std::move(b))) return; }
        if ((c.op == "Cast" || c.op == "Mul" || c.op == "Reshape") && !c.out.empty()) next = c.out[0];
      }
      v = next;
    }
  }

  // scale of a MatMulInteger / ConvInteger weight when the "<base>_scale" naming is absent: the integer result is
  // cast and multiplied by (a_scale * b_scale); b_scale is the constant operand of that product
  FT find_scale_downstream(const ONode& n) {
    FT none;
    if (n.out.empty()) return none;
    std::string v = n.out[0];
    for (int hop = 0; hop < 3 && !v.empty(); hop++) {
      auto range = g_.consumers.equal_range(v);
      std::string next;
      for (auto it = range.first; it != range.second; ++it) {
        const ONode& c = g_.nodes[it->second];
        if (c.op == "Cast" && !c.out.empty()) next = c.out[0];
        if (c.op == "Mul") {
          for (auto& in : c.in) {
            if (in == v) continue;
            FT s = value(in);
            if (s.ok) return s;
            auto pit = g_.producer.find(in);
            if (pit != g_.producer.end() && g_.nodes[pit->second].op == "Mul")
              for (auto& in2 : g_.nodes[pit->second].in) { FT s2 = value(in2); if (s2.ok) return s2; }
          }
        }
      }
      v = next;
    }
    return none;
  }

  void place_lstm(const std::string& mod, const FT& W, const FT& R, const FT& Bv) {
    // W [D,4H,I], R [D,4H,H], B [D,8H] in ONNX gate order i,o,f,c -> torch i,f,g,o
    if (W.dims.size() != 3 || R.dims.size() != 3) return;
    const int D = (int)W.dims[0], H4 = (int)W.dims[1], I = (int)W.dims[2], H = H4 / 4;
    if ((int)R.dims[0] != D || (int)R.dims[1] != H4 || (int)R.dims[2] != H) return;
    const int src_of[4] = {0, 2, 3, 1};     // torch gate k (i,f,g,o) sits at ONNX position src_of[k] (i,o,f,c)
    for (int d = 0; d < D && d < 2; d++) {
      const std::string sfx = std::string("_l0") + (d == 1 ? "_reverse" : "");
      FT wih, whh, bih, bhh;
      wih.ok = whh.ok = bih.ok = bhh.ok = true;
      wih.v.resize((size_t)H4 * I); whh.v.resize((size_t)H4 * H); bih.v.assign(H4, 0.f); bhh.v.assign(H4, 0.f);
      for (int k = 0; k < 4; k++)
        for (int r = 0; r < H; r++) {
          const size_t dst = (size_t)k * H + r, src = (size_t)src_of[k] * H + r;
          memcpy(&wih.v[dst * I], &W.v[((size_t)d * H4 + src) * I], (size_t)I * 4);
          memcpy(&whh.v[dst * H], &R.v[((size_t)d * H4 + src) * H], (size_t)H * 4);
          if (Bv.ok && Bv.v.size() == (size_t)D * 2 * H4) {
            bih.v[dst] = Bv.v[(size_t)d * 2 * H4 + src];
            bhh.v[dst] = Bv.v[(size_t)d * 2 * H4 + H4 + src];
          }
        }
      place(mod + ".weight_ih" + sfx, std::move(wih)); place(mod + ".weight_hh" + sfx, std::move(whh));
      place(mod + ".bias_ih" + sfx, std::move(bih)); place(mod + ".bias_hh" + sfx, std::move(bhh));
    }
  }

  // [D, A, B] -> [D, B, A]
  static FT swap_last2(const FT& x) {
    FT o;
    o.ok = true;
    const size_t D = (size_t)x.dims[0], A = (size_t)x.dims[1], B = (size_t)x.dims[2];
    o.dims = {x.dims[0], x.dims[2], x.dims[1]};
    o.v.resize(x.v.size());
    for (size_t d = 0; d < D; d++)
      for (size_t a = 0; a < A; a++)
        for (size_t b = 0; b < B; b++) o.v[(d * B + b) * A + a] = x.v[(d * A + a) * B + b];
    return o;
  }

  void by_graph() {
    for (const ONode& n : g_.nodes) {
      const std::string& op = n.op;
      const bool conv = op == "Conv" || op == "ConvTranspose", convi = op == "ConvInteger";
      const bool mm = op == "MatMul", mmi = op == "MatMulInteger", gemm = op == "Gemm", nbits = op == "MatMulNBits";
      const bool lstm = op == "LSTM", qlstm = op == "DynamicQuantizeLSTM";
      if (!(conv || convi || mm || mmi || gemm || nbits || lstm || qlstm)) continue;
      const std::string mod = module_of(n);
      if (mod.empty()) continue;
      if (lstm && n.in.size() >= 3) {
        FT W = value(n.in[1]), R = value(n.in[2]), B;
        if (n.in.size() > 3 && !n.in[3].empty()) B = value(n.in[3]);
        if (W.ok && R.ok) place_lstm(mod, W, R, B);
        continue;
      }
      if (qlstm && n.in.size() >= 12) {
        // com.microsoft DynamicQuantizeLSTM: W [D, I, 4H] and R [D, H, 4H] quantised (transposed), scales / zero
        // points per direction at inputs 8..11
        FT W = value(n.in[1]), R = value(n.in[2]), B;
        if (!n.in[3].empty()) B = value(n.in[3]);
        FT ws = value(n.in[8]), wz = value(n.in[9]), rs = value(n.in[10]), rz = value(n.in[11]);
        if (W.ok && R.ok && ws.ok && rs.ok && W.dims.size() == 3 && R.dims.size() == 3) {
          FT Wd = dequant(W, ws, wz, ws.v.size() > 1 ? 0 : 0), Rd = dequant(R, rs, rz, 0);
          place_lstm(mod, swap_last2(Wd), swap_last2(Rd), B);
        }
        continue;
      }
      if (!wanted(mod + ".weight")) continue;
      if (conv && n.in.size() >= 2) {
        FT w = value(n.in[1]);
        if (w.ok) place(mod + ".weight", std::move(w));
        if (n.in.size() > 2 && !out_.has(mod + ".bias") && wanted(mod + ".bias")) { FT b = value(n.in[2]); if (b.ok) place(mod + ".bias", std::move(b)); }
      } else if (convi && n.in.size() >= 2) {
        FT q = value(n.in[1]);
        if (!q.ok) continue;
        if (g_.init.count(n.in[1]) && is_quant_type(g_.init.at(n.in[1]).dtype) && strip_suffix(n.in[1], "_quantized").empty()) {
          FT zp; if (n.in.size() > 3 && !n.in[3].empty()) zp = value(n.in[3]);
          FT sc = find_scale_downstream(n);
          if (!sc.ok) continue;
          q = dequant(q, sc, zp, 0);
        }
        place(mod + ".weight", std::move(q));
      } else if (mm && n.in.size() >= 2) {
        FT w = value(n.in[1]);
        if (!w.ok) w = value(n.in[0]);
        if (!w.ok || w.dims.size() != 2) continue;
        place(mod + ".weight", transpose2(w));       // [K, N] -> torch Linear [N, K]
        place_linear_bias(n, mod);
      } else if (mmi && n.in.size() >= 2) {
        FT q = value(n.in[1]);
        if (!q.ok || q.dims.size() != 2) continue;
        if (g_.init.count(n.in[1]) && is_quant_type(g_.init.at(n.in[1]).dtype) && strip_suffix(n.in[1], "_quantized").empty()) {
          FT zp; if (n.in.size() > 3 && !n.in[3].empty()) zp = value(n.in[3]);
          FT sc = find_scale_downstream(n);
          if (!sc.ok) continue;
          q = dequant(q, sc, zp, 1);
        }
        place(mod + ".weight", transpose2(q));
        place_linear_bias(n, mod);
      } else if (gemm && n.in.size() >= 2) {
        FT w = value(n.in[1]);
        if (!w.ok || w.dims.size() != 2) continue;
        const bool transB = n.ints.count("transB") && n.ints.at("transB") != 0;
        place(mod + ".weight", transB ? std::move(w) : transpose2(w));
        if (n.in.size() > 2 && !out_.has(mod + ".bias") && wanted(mod + ".bias")) { FT b = value(n.in[2]); if (b.ok) place(mod + ".bias", std::move(b)); }
      } else if (nbits && n.in.size() >= 3) {
        place_nbits(n, mod);
        place_linear_bias(n, mod);
      }
    }
  }

  // com.microsoft MatMulNBits (bits = 4): B [N][K/bs][bs/2] uint8 (low nibble first), scales [N * K/bs],
  // zero_points packed 4-bit [N][ceil(K/bs/2)] (default 8) -> torch Linear weight [N][K]
  void place_nbits(const ONode& n, const std::string& mod) {
    const long long K = n.ints.count("K") ? n.ints.at("K") : 0, N = n.ints.count("N") ? n.ints.at("N") : 0;
    const long long bits = n.ints.count("bits") ? n.ints.at("bits") : 4, bs = n.ints.count("block_size") ? n.ints.at("block_size") : 32;
    if (bits != 4 || K <= 0 || N <= 0 || bs < 16 || (bs & (bs - 1))) { notes_ += "MatMulNBits with bits != 4 or odd block size at " + mod; return; }
    auto bi = g_.init.find(n.in[1]);
    if (bi == g_.init.end() || !bi->second.raw.p) return;
    FT sc = value(n.in[2]);
    if (!sc.ok) return;
    const long long nblk = (K + bs - 1) / bs;
    if ((long long)bi->second.raw.n != N * nblk * (bs / 2) || (long long)sc.v.size() != N * nblk) return;
    const uint8_t* zp = nullptr;
    if (n.in.size() > 3 && !n.in[3].empty()) {
      auto zi = g_.init.find(n.in[3]);
      if (zi != g_.init.end() && zi->second.raw.p && zi->second.dtype == DT_U8 && (long long)zi->second.raw.n == N * ((nblk + 1) / 2)) {
        zp = zi->second.raw.p;
        used_.insert(n.in[3]);
      } else return;      // float zero points etc.: not handled
    }
    used_.insert(n.in[1]);
    FT w;
    w.ok = true; w.dims = {N, K}; w.v.resize((size_t)N * K);
    const uint8_t* B = bi->second.raw.p;
    for (long long r = 0; r < N; r++)
      for (long long b = 0; b < nblk; b++) {
        const float s = sc.v[(size_t)(r * nblk + b)];
        int z = 8;
        if (zp) { const uint8_t zz = zp[r * ((nblk + 1) / 2) + b / 2]; z = (b & 1) ? (zz >> 4) : (zz & 15); }
        for (long long j = 0; j < bs && b * bs + j < K; j++) {
          const uint8_t byte = B[(r * nblk + b) * (bs / 2) + j / 2];
          const int q = (j & 1) ? (byte >> 4) : (byte & 15);
          w.v[(size_t)(r * K + b * bs + j)] = (float)(q - z) * s;
        }
      }
    place(mod + ".weight", std::move(w));
  }

  void fold_weight_norm() {
    for (auto& name : out_.names()) {
      std::string base, vname;
      if (!(base = strip_suffix(name, ".weight_g")).empty()) vname = base + ".weight_v";
      else if (!(base = strip_suffix(name, ".parametrizations.weight.original0")).empty()) vname = base + ".parametrizations.weight.original1";
      else continue;
      const std::string wname = base + ".weight";
      if (!wanted(wname) || out_.has(wname) || !out_.has(vname)) continue;
      const HostTensor& g = out_.get(name);
      const HostTensor& v = out_.get(vname);
      if (v.shape.empty() || g.numel != (size_t)v.shape[0]) continue;
      const size_t rows = (size_t)v.shape[0], inner = v.numel / rows;
      FT w;
      w.ok = true; w.v.resize(v.numel);
      for (size_t r = 0; r < rows; r++) {
        double ss = 0;
        for (size_t i = 0; i < inner; i++) ss += (double)v.data[r * inner + i] * v.data[r * inner + i];
        const double k = (double)g.data[r] / std::sqrt(ss);
        for (size_t i = 0; i < inner; i++) w.v[r * inner + i] = (float)(v.data[r * inner + i] * k);
      }
      place(wname, std::move(w));
    }
  }

  Graph g_;
  WeightFile& out_;
  std::map<std::string, std::vector<int>> need_;
  std::vector<std::string> order_;
  std::set<std::string> used_;
  std::string notes_;
};

}  // namespace

void load_onnx_weights(const std::string& path, const std::vector<char>& bytes, WeightFile& out) {
  Reader r(parse_model(bytes, path), out);
  r.run(path);
}

}  // namespace kkx
