// compute_matches (putative-matching stage) -- the reference's apps/compute_matches/compute_matches.cpp:47-248 with
// the CUDA back-end and a multi-GPU pair scheduler.  Same flags, defaults, files and resume rules:
//
//   -i/--imadir <dir>  -o/--outdir <dir>  [-r/--distratio 0.6]  [-s/--isZoom 0]  [-p/--contrastThreshold 0.04]
//   [-g/--geometricModel f|e|h]            (compute_matches.cpp:51-68; spellings of cmd_line.h:80-97:
//                                           "-r 0.8", "-r0.8", "--distratio 0.8", "--distratio=0.8")
//   --gpus N                               new: shard the i<j pair list over N GPUs (default: all visible)
//
// Reads <outdir>/lists.txt (image_list_io_helper.h:69-193; only the file names are used here), for every image
// <outdir>/<basename>.feat (feature.h:117-133) and .desc (descriptor.h:160-181; 8-byte or 4-byte count), and writes
// <outdir>/matches.putative.txt byte-identical to PairedIndexedMatchToStream (indexed_match_utils.h:22-38) of the
// reference's BRUTE-FORCE matcher (the typedef at compute_matches.cpp:226-227; the shipped default :222-223 is the
// approximate, randomised FLANN kd-tree and is deliberately not reproduced).
// Resume rule kept: if matches.putative.txt exists, matching is skipped (compute_matches.cpp:230-234).
// Out of scope of this build: SIFT extraction (the stage before: .feat/.desc must exist) and
// the AC-RANSAC geometric filter (the stage after: run the reference's own binary on the exported file).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "mvgcuda.h"

namespace {

struct Options {
  std::string imadir, outdir, geometric_model = "f";
  float dist_ratio = 0.6f;   // compute_matches.cpp:55
  bool is_zoom = false;
  float contrast_threshold = 0.04f;
  int gpus = -1;
};

template <typename T>
bool parse_value(const std::string& tok, T& out) {  // cmd_line.h:110-113: must consume the whole token
  std::istringstream ss(tok);
  ss >> out;
  char c;
  return !ss.fail() && !(ss >> c);
}
bool parse_value(const std::string& tok, std::string& out) { out = tok; return true; }

// one option in the four spellings of cmd_line.h:80-97; returns 1 if consumed (and advances i), 0 if not this option, -1 on error
template <typename T>
int take(char c, const char* longname, T& dst, std::vector<std::string>& a, size_t& i) {
  const std::string& s = a[i];
  std::string val;
  bool have = false;
  const std::string l = std::string("--") + longname;
  if (s.size() >= 2 && s[0] == '-' && s[1] == c && s[1] != '-') {
    if (s.size() > 2) { val = s.substr(2); have = true; }
    else if (i + 1 < a.size()) { val = a[++i]; have = true; }
    else return -1;
  } else if (s == l) {
    if (i + 1 < a.size()) { val = a[++i]; have = true; } else return -1;
  } else if (s.compare(0, l.size() + 1, l + "=") == 0) {
    val = s.substr(l.size() + 1); have = true;
  }
  if (!have) return 0;
  return parse_value(val, dst) ? 1 : -1;
}

bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
bool dir_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

std::string basename_part(const std::string& f) {  // file_system.h basename_part: strip directory and extension
  size_t s = f.find_last_of("/\\");
  std::string b = s == std::string::npos ? f : f.substr(s + 1);
  size_t d = b.find_last_of('.');
  return d == std::string::npos ? b : b.substr(0, d);
}

bool load_list(const std::string& path, std::vector<std::string>& names) {
  std::ifstream in(path.c_str());
  if (!in.is_open()) return false;
  std::string line;
  while (std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    if (line.empty()) continue;
    names.push_back(line.substr(0, line.find(';')));
  }
  return true;
}

bool load_desc(const std::string& path, std::vector<uint8_t>& out, int& rows) {
  std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
  rows = 0;
  out.clear();
  if (!f.is_open()) return true;  // reference: missing file -> empty set, "ok" (descriptor.h:168-180)
  const std::streamoff size = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> raw((size_t)size);
  if (size) f.read(reinterpret_cast<char*>(raw.data()), size);
  for (int hdr : {8, 4}) {  // sizeof(size_t) of the writer: 8 on Linux x86-64, 4 in the shipped data/et files
    if (size < hdr) continue;
    uint64_t n = 0;
    memcpy(&n, raw.data(), hdr);
    if ((uint64_t)size == (uint64_t)hdr + n * MVGCUDA_DIM) {
      out.assign(raw.begin() + hdr, raw.end());
      rows = (int)n;
      return true;
    }
  }
  return size == 0;
}

bool load_feat_xy(const std::string& path, std::vector<float>& xy) {
  xy.clear();
  std::ifstream f(path.c_str());
  if (!f.is_open()) return true;
  float x, y, s, o;
  while (f >> x >> y >> s >> o) { xy.push_back(x); xy.push_back(y); }
  return !f.bad();
}

// decimal text of a non-negative integer, appended to a buffer (the export is hundreds of MB of small integers)
inline void append_uint(std::string& out, unsigned long long v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) out.push_back(tmp[--n]);
}

}  // namespace

int main(int argc, char** argv) {
  Options opt;
  std::vector<std::string> a(argv + 1, argv + argc);
  for (size_t i = 0; i < a.size(); ++i) {
    int r = 0;
    std::string g;
    if ((r = take('i', "imadir", opt.imadir, a, i))) {}
    else if ((r = take('o', "outdir", opt.outdir, a, i))) {}
    else if ((r = take('r', "distratio", opt.dist_ratio, a, i))) {}
    else if ((r = take('s', "isZoom", opt.is_zoom, a, i))) {}
    else if ((r = take('p', "contrastThreshold", opt.contrast_threshold, a, i))) {}
    else if ((r = take('g', "geometricModel", opt.geometric_model, a, i))) {}
    else if (a[i] == "--gpus" && i + 1 < a.size()) { r = parse_value(a[++i], opt.gpus) ? 1 : -1; }
    else if (a[i].compare(0, 7, "--gpus=") == 0) { r = parse_value(a[i].substr(7), opt.gpus) ? 1 : -1; }
    else { std::cerr << "Unrecognized option " << a[i] << std::endl; r = -1; }
    if (r < 0) {
      std::cerr << "Usage: " << argv[0] << " -i|--imadir path -o|--outdir path [-r|--distratio 0.6] [-s|--isZoom 0] "
                << "[-p|--contrastThreshold 0.04] [-g|--geometricModel f|e|h] [--gpus N]" << std::endl;
      return EXIT_FAILURE;
    }
  }
  if (opt.outdir.empty()) { std::cerr << "\nIt is an invalid output directory" << std::endl; return EXIT_FAILURE; }
  std::cout << " You called : \n" << argv[0] << "\n--imadir " << opt.imadir << "\n--outdir " << opt.outdir << "\n--distratio "
            << opt.dist_ratio << "\n--geometricModel " << opt.geometric_model << std::endl;
  if (!dir_exists(opt.outdir)) { std::cerr << "output directory " << opt.outdir << " does not exist" << std::endl; return EXIT_FAILURE; }

  std::vector<std::string> names;
  if (!load_list(opt.outdir + "/lists.txt", names)) {
    std::cerr << "\nEmpty or invalid image list: " << opt.outdir << "/lists.txt" << std::endl;
    return EXIT_FAILURE;
  }
  const std::string putative = opt.outdir + "/matches.putative.txt";
  if (file_exists(putative)) {  // compute_matches.cpp:230-234
    std::cout << "\nPREVIOUS RESULTS LOADED: " << putative << " exists, putative matching skipped" << std::endl;
    return EXIT_SUCCESS;
  }

  const int n = (int)names.size();
  std::vector<std::vector<uint8_t> > desc(n);
  std::vector<std::vector<float> > xy(n);
  std::vector<int32_t> rows(n, 0);
  auto t0 = std::chrono::steady_clock::now();
  {
    // one image per task over the host cores: the .feat text parse dominates (4 floats per feature)
    std::atomic<int> next(0), first_missing(n), first_bad(n);
    auto load = [&]() {
      for (int i = next++; i < n; i = next++) {
        const std::string base = opt.outdir + "/" + basename_part(names[i]);
        if (!file_exists(base + ".feat") || !file_exists(base + ".desc")) {
          int cur = first_missing.load();
          while (i < cur && !first_missing.compare_exchange_weak(cur, i)) {}
          continue;
        }
        int drows = 0;
        if (!load_desc(base + ".desc", desc[i], drows) || !load_feat_xy(base + ".feat", xy[i])) {
          int cur = first_bad.load();
          while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
          continue;
        }
        rows[i] = (int32_t)std::min<size_t>(xy[i].size() / 2, (size_t)drows);  // row count from the features (matcher_all_in_memory.h:80)
        xy[i].resize(2 * (size_t)rows[i]);
      }
    };
    const int n_threads = (int)std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)std::max(n, 1)));
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(load);
    for (auto& t : pool) t.join();
    if (first_missing.load() < n) {
      const std::string base = opt.outdir + "/" + basename_part(names[first_missing.load()]);
      std::cerr << "missing " << base << ".feat/.desc: SIFT extraction (compute_matches.cpp:188-216) is outside this build; "
                << "run the reference's extraction stage first" << std::endl;
      return EXIT_FAILURE;
    }
    if (first_bad.load() < n) {
      std::cerr << "cannot parse " << opt.outdir << "/" << basename_part(names[first_bad.load()]) << ".feat/.desc" << std::endl;
      return EXIT_FAILURE;
    }
  }
  std::vector<int32_t> pairs;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) { pairs.push_back(i); pairs.push_back(j); }
  const int64_t n_pairs = (int64_t)pairs.size() / 2;

  int gpus = opt.gpus > 0 ? opt.gpus : mvgcuda_device_count();
  if (gpus < 1) { std::cerr << "no sm_100 CUDA device visible (there is no CPU fallback)" << std::endl; return EXIT_FAILURE; }
  gpus = (int)std::max<int64_t>(1, std::min<int64_t>(gpus, n_pairs));
  std::cout << std::endl << " - PUTATIVE MATCHES - " << n << " images, " << n_pairs << " pairs, " << gpus << " GPU(s)" << std::endl;

  const float ratio_sq = opt.dist_ratio * opt.dist_ratio;  // Square(float), numeric.h:108-111: rounded in fp32
  std::vector<int64_t> bounds(gpus + 1, n_pairs);
  {
    std::vector<double> csum(n_pairs + 1, 0.0);
    for (int64_t p = 0; p < n_pairs; ++p)
      csum[p + 1] = csum[p] + std::max(1.0, (double)rows[pairs[2 * p]] * (double)rows[pairs[2 * p + 1]]);
    bounds[0] = 0;
    for (int g = 1; g < gpus; ++g) bounds[g] = std::lower_bound(csum.begin(), csum.end(), csum[n_pairs] * g / gpus) - csum.begin();
    for (int g = 1; g <= gpus; ++g) bounds[g] = std::max(bounds[g], bounds[g - 1]);
  }
  std::vector<const uint8_t*> dptr(n);
  std::vector<const float*> fptr(n);
  for (int i = 0; i < n; ++i) { dptr[i] = rows[i] ? desc[i].data() : NULL; fptr[i] = xy[i].data(); }

  struct Shard { std::vector<int32_t> counts; std::vector<int32_t> matches; std::string error; float gpu_ms = 0; };
  std::vector<Shard> shards(gpus);
  auto t1 = std::chrono::steady_clock::now();
  // GPU 0 reads the collection over PCIe once; every other GPU takes a replica of its arena device to device
  // (NVLink / NVSwitch), then all match their shard of the pair list concurrently -- no collective afterwards
  std::vector<mvgcuda_ctx*> ctxs(gpus, (mvgcuda_ctx*)NULL);
  auto destroy_all = [&]() { for (mvgcuda_ctx* c : ctxs) if (c) mvgcuda_destroy(c); };
  if (mvgcuda_create(0, &ctxs[0]) != MVGCUDA_OK) { std::cerr << "GPU 0: " << mvgcuda_last_error(NULL) << std::endl; return EXIT_FAILURE; }
  if (mvgcuda_upload_images(ctxs[0], n, dptr.data(), rows.data(), 0) != MVGCUDA_OK ||
      mvgcuda_set_features(ctxs[0], n, fptr.data(), rows.data()) != MVGCUDA_OK) {
    std::cerr << "GPU 0: " << mvgcuda_last_error(ctxs[0]) << std::endl;
    destroy_all();
    return EXIT_FAILURE;
  }
  auto work = [&](int g) {
    Shard& S = shards[g];
    if (g > 0) {
      if (mvgcuda_create(g, &ctxs[g]) != MVGCUDA_OK) { S.error = mvgcuda_last_error(NULL); return; }
      if (mvgcuda_clone_images(ctxs[g], ctxs[0]) != MVGCUDA_OK) { S.error = mvgcuda_last_error(ctxs[g]); return; }
    }
    mvgcuda_ctx* ctx = ctxs[g];
    const int64_t b = bounds[g], e = bounds[g + 1];
    mvgcuda_pair_matches pm;
    if (mvgcuda_match_collection(ctx, e - b, pairs.data() + 2 * b, ratio_sq, 0, &pm) != MVGCUDA_OK) {
      S.error = mvgcuda_last_error(ctx);
    } else {
      S.counts.assign(pm.counts, pm.counts + (e - b));
      S.matches.assign(pm.matches, pm.matches + 2 * pm.offsets[e - b]);
      S.gpu_ms = pm.gpu_ms;
    }
  };
  {
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g) th.emplace_back(work, g);
    for (auto& t : th) t.join();
  }
  destroy_all();  // only after every replica has been taken
  auto t2 = std::chrono::steady_clock::now();
  for (int g = 0; g < gpus; ++g)
    if (!shards[g].error.empty()) { std::cerr << "GPU " << g << ": " << shards[g].error << std::endl; return EXIT_FAILURE; }

  // export: pairs are generated in lexicographic (i,j) order == std::map iteration order
  FILE* f = fopen(putative.c_str(), "wb");
  if (!f) { std::cerr << "cannot write " << putative << std::endl; return EXIT_FAILURE; }
  long long total = 0;
  {
    // text of every shard formatted concurrently, written in shard (== pair) order
    std::vector<std::string> text(gpus);
    std::vector<long long> shard_total(gpus, 0);
    auto format = [&](int g) {
      const Shard& S = shards[g];
      std::string& out = text[g];
      out.reserve(S.matches.size() * 6 + (size_t)(bounds[g + 1] - bounds[g]) * 16 + 64);
      size_t off = 0;
      for (int64_t p = bounds[g]; p < bounds[g + 1]; ++p) {
        const int c = S.counts[p - bounds[g]];
        append_uint(out, (unsigned)pairs[2 * p]); out.push_back(' ');
        append_uint(out, (unsigned)pairs[2 * p + 1]); out.push_back('\n');
        append_uint(out, (unsigned)c); out.push_back('\n');
        for (int k = 0; k < c; ++k, ++off) {
          append_uint(out, (unsigned)S.matches[2 * off]); out.push_back(' ');
          append_uint(out, (unsigned)S.matches[2 * off + 1]); out.push_back('\n');
        }
        shard_total[g] += c;
      }
    };
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g) th.emplace_back(format, g);
    for (auto& t : th) t.join();
    for (int g = 0; g < gpus; ++g) {
      fwrite(text[g].data(), 1, text[g].size(), f);
      total += shard_total[g];
    }
  }
  if (fclose(f) != 0) { std::cerr << "short write to " << putative << std::endl; return EXIT_FAILURE; }
  auto t3 = std::chrono::steady_clock::now();
  auto ms = [](std::chrono::steady_clock::time_point a_, std::chrono::steady_clock::time_point b_) {
    return std::chrono::duration<double, std::milli>(b_ - a_).count();
  };
  std::cout << "loaded in " << ms(t0, t1) << " ms; matched " << n_pairs << " pairs (" << total << " putative matches) in " << ms(t1, t2)
            << " ms on " << gpus << " GPU(s) (" << (n_pairs / std::max(1e-9, ms(t1, t2) * 1e-3)) << " pairs/s incl. upload + host de-dup); exported in "
            << ms(t2, t3) << " ms -> " << putative << std::endl;
  std::cout << "geometric filtering (-g " << opt.geometric_model << ") is the next stage and is not part of this build" << std::endl;
  return EXIT_SUCCESS;
}
