// compute_matches (putative-matching stage) -- the reference's apps/compute_matches/compute_matches.cpp:47-248 with
// the CUDA back-end and a streaming multi-GPU pair scheduler.  Same flags, defaults, files and resume rules:
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
// Resume rule kept: if matches.putative.txt exists it is imported (pairedIndexedMatchImport, indexed_match_utils.h:48-73)
// and matching is skipped (compute_matches.cpp:230-234).
//
// Scheduler (what replaces MatcherAllInMemory::LoadData + Match, matcher_all_in_memory.h:44-141):
//   1. the .desc headers are scanned for the row counts, then one context per GPU is created, all concurrently;
//   2. loader threads (all host cores) parse one image each -- binary .desc straight into page-locked staging, text .feat
//      -- and hand it to EVERY GPU at once (mvgcuda_stream_image: asynchronous copy + norms kernel), so files are still
//      being parsed while earlier images travel; each GPU keeps a full replica (<= ~1 GB), there is no collective;
//   3. the i < j pair list is cut into contiguous, cost-balanced (rows_i * rows_j) shards, one per GPU; every GPU runs
//      rows 7-13 of the path on its shard (mvgcuda_match_collection; batches pipelined inside the library);
//   4. the text of every shard is formatted concurrently and written in pair order.
//   5. -g f / -g h: the AC-RANSAC filter (compute_matches.cpp:250-318, GeometricFilter_FMatrix_AC(4.0) resp.
//      GeometricFilter_HMatrix_AC(4.0)) runs on GPU 0 over the putative matches (fresh or imported) and writes
//      <outdir>/matches.f.txt resp. matches.h.txt; the reference's rand() stream is never seeded (== srand(1)) and is
//      consumed pair after pair in map order -- reproduced.
// Out of scope of this build: SIFT extraction (the stage before: .feat/.desc must exist) and the
// essential variant of the filter (-g e, the reference's default, needs K.txt and the 5-point solver): it stops after the
// putative stage with a message.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <iostream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "mvgcuda.h"

namespace {

struct Options {
  std::string imadir, outdir, geometric_model = "e";  // compute_matches.cpp:58
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

// lists.txt: "name;width;height[;focal;...]" (image_list_io_helper.h:69-193); sizes gets width, height per image (0 when absent)
bool load_list(const std::string& path, std::vector<std::string>& names, std::vector<int32_t>& sizes) {
  std::ifstream in(path.c_str());
  if (!in.is_open()) return false;
  std::string line;
  while (std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    if (line.empty()) continue;
    std::vector<std::string> f;
    size_t b = 0;
    for (;;) {
      const size_t e = line.find(';', b);
      f.push_back(line.substr(b, e == std::string::npos ? std::string::npos : e - b));
      if (e == std::string::npos) break;
      b = e + 1;
    }
    names.push_back(f[0]);
    int w = 0, h = 0;
    if (f.size() >= 3) { w = atoi(f[1].c_str()); h = atoi(f[2].c_str()); }
    sizes.push_back(w);
    sizes.push_back(h);
  }
  return true;
}

// "x y scale orientation" per feature (feature.h:117-133), parsed from a buffer that holds the whole file: the same token
// rule as  in >> x >> y >> scale >> orientation  (whitespace-separated, stop at the first token that is not a number), an
// order of magnitude faster than the stream extractors (std::from_chars rounds like strtof).  xy receives the first
// `cap` (x, y); returns the number of complete features in the file.
size_t parse_feat_buffer(const char* p, const char* end, float* xy, size_t cap) {
  size_t got = 0;
  for (;;) {
    float v[4];
    int k = 0;
    for (; k < 4; ++k) {
      while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t' || *p == '\f' || *p == '\v')) ++p;
      if (p < end && *p == '+') ++p;
      const std::from_chars_result r = std::from_chars(p, end, v[k]);
      if (r.ec != std::errc() || r.ptr == p) break;
      p = r.ptr;
    }
    if (k < 4) break;
    if (got < cap) { xy[2 * got] = v[0]; xy[2 * got + 1] = v[1]; }
    ++got;
  }
  return got;
}

// decimal text of a non-negative integer, appended to a buffer (the export is hundreds of MB of small integers)
inline void append_uint(std::string& out, unsigned long long v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) out.push_back(tmp[--n]);
}

// rows of a .desc file from its header and size, without reading the payload (0 for a missing / empty file; -1 malformed)
int desc_rows_from_header(const std::string& path, int& header_bytes) {
  header_bytes = 0;
  std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
  if (!f.is_open()) return 0;
  const std::streamoff size = f.tellg();
  if (size == 0) return 0;
  f.seekg(0);
  unsigned char h[8] = {0};
  f.read(reinterpret_cast<char*>(h), std::min<std::streamoff>(8, size));
  for (int hdr : {8, 4}) {  // sizeof(size_t) of the writer: 8 on Linux x86-64, 4 in the shipped data/et files
    if (size < hdr) continue;
    uint64_t n = 0;
    memcpy(&n, h, hdr);
    if ((uint64_t)size == (uint64_t)hdr + n * MVGCUDA_DIM) { header_bytes = hdr; return (int)n; }
  }
  return -1;
}

// pairedIndexedMatchImport (indexed_match_utils.h:48-73): "i j\ncount\n" + count x "_i _j\n" until the first token that
// does not parse; a key seen twice keeps its LAST block (operator[] assignment).  Returns false if the file cannot be opened.
struct ImportedMatches {
  std::vector<std::pair<std::pair<size_t, size_t>, std::vector<std::pair<int, int> > > > blocks;  // file order
  size_t pairs = 0, matches = 0;  // after the map semantics (distinct keys, last block wins)
};
bool import_matches(const std::string& path, ImportedMatches& out) {
  std::ifstream in(path.c_str());
  if (!in.is_open()) {
    std::cout << std::endl << "ERROR IndexedMatchesUtils::import(...)" << std::endl << "with : " << path << std::endl;
    return false;
  }
  size_t l, r, number;
  while (in >> l >> r >> number) {
    std::vector<std::pair<int, int> > m(number);
    for (size_t k = 0; k < number; ++k) in >> m[k].first >> m[k].second;
    out.blocks.push_back(std::make_pair(std::make_pair(l, r), m));
  }
  std::vector<size_t> order(out.blocks.size());
  for (size_t k = 0; k < order.size(); ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return out.blocks[a].first < out.blocks[b].first; });
  for (size_t k = 0; k < order.size(); ++k) {
    if (k + 1 < order.size() && out.blocks[order[k + 1]].first == out.blocks[order[k]].first) continue;  // a later block of the key wins
    ++out.pairs;
    out.matches += out.blocks[order[k]].second.size();
  }
  return true;
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
  // cuInit enumerates every visible GPU (5-6 s on a box with eight 180 GB devices, 0.3 s with one): with an explicit
  // --gpus N and no CUDA_VISIBLE_DEVICES of the caller's, only the first N devices are made visible -- before the first
  // CUDA call of the process.
  if (opt.gpus > 0 && getenv("CUDA_VISIBLE_DEVICES") == NULL) {
    std::string vis;
    for (int g = 0; g < opt.gpus; ++g) vis += (g ? "," : "") + std::to_string(g);
    setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
  }
  std::cout << " You called : \n" << argv[0] << "\n--imadir " << opt.imadir << "\n--outdir " << opt.outdir << "\n--distratio "
            << opt.dist_ratio << "\n--geometricModel " << opt.geometric_model << std::endl;
  if (!dir_exists(opt.outdir)) { std::cerr << "output directory " << opt.outdir << " does not exist" << std::endl; return EXIT_FAILURE; }
  // the model is the first character of -g, either case; anything else ends the run (compute_matches.cpp:99-116)
  char model_char = opt.geometric_model.empty() ? '\0' : opt.geometric_model[0];
  if (model_char >= 'A' && model_char <= 'Z') model_char = static_cast<char>(model_char - 'A' + 'a');
  if (model_char != 'f' && model_char != 'e' && model_char != 'h') {
    std::cerr << "Unknown geometric model" << std::endl;
    return EXIT_FAILURE;
  }

  std::vector<std::string> names;
  std::vector<int32_t> image_sizes;
  if (!load_list(opt.outdir + "/lists.txt", names, image_sizes)) {
    std::cerr << "\nEmpty or invalid image list: " << opt.outdir << "/lists.txt" << std::endl;
    return EXIT_FAILURE;
  }
  const std::string putative = opt.outdir + "/matches.putative.txt";
  const bool filter_f = model_char == 'f' || model_char == 'h';  // the models built on the GPU
  const char geo_model = model_char == 'h' ? 'h' : 'f';
  const bool resumed = file_exists(putative);
  ImportedMatches im;
  if (resumed) {  // compute_matches.cpp:230-234
    if (!import_matches(putative, im)) return EXIT_FAILURE;
    std::cout << std::endl << "PUTATIVE MATCHES -- PREVIOUS RESULTS LOADED" << std::endl
              << im.pairs << " pairs, " << im.matches << " putative matches imported from " << putative << "; matching skipped" << std::endl;
    if (!filter_f) {
      std::cout << "geometric filtering with -g " << opt.geometric_model << " is not part of this build (only -g f and -g h are)" << std::endl;
      return EXIT_SUCCESS;
    }
  }

  typedef std::chrono::steady_clock Clock;
  auto ms = [](Clock::time_point a_, Clock::time_point b_) { return std::chrono::duration<double, std::milli>(b_ - a_).count(); };
  const Clock::time_point t_begin = Clock::now();
  const int n = (int)names.size();
  const int64_t n_pairs = (int64_t)n * (n - 1) / 2;
  // ---- 1. header scan (row counts), then one context per GPU, created concurrently
  std::vector<int32_t> rows(n, 0);
  std::vector<int> hdr_bytes(n, 0);
  for (int i = 0; i < n; ++i) {
    const std::string base = opt.outdir + "/" + basename_part(names[i]);
    if (!file_exists(base + ".feat") || !file_exists(base + ".desc")) {
      std::cerr << "missing " << base << ".feat/.desc: SIFT extraction (compute_matches.cpp:188-216) is outside this build; "
                << "run the reference's extraction stage first" << std::endl;
      return EXIT_FAILURE;
    }
    const int r = desc_rows_from_header(base + ".desc", hdr_bytes[i]);
    if (r < 0) { std::cerr << "cannot parse " << base << ".desc" << std::endl; return EXIT_FAILURE; }
    rows[i] = r;
  }
  const int n_devices = mvgcuda_device_count();
  if (n_devices < 1) { std::cerr << "no sm_100 CUDA device visible (there is no CPU fallback)" << std::endl; return EXIT_FAILURE; }
  int gpus = opt.gpus > 0 ? opt.gpus : n_devices;  // more shards than devices: they share devices round-robin
  gpus = (int)std::max<int64_t>(1, std::min<int64_t>(gpus, std::max<int64_t>(n_pairs, 1)));
  if (resumed) gpus = 1;  // only the filter runs, and the chain of its rand() stream is walked by one GPU
  std::vector<mvgcuda_ctx*> ctxs(gpus, (mvgcuda_ctx*)NULL);
  std::vector<std::string> ctx_error(gpus);
  auto destroy_all = [&]() { for (mvgcuda_ctx* c : ctxs) if (c) mvgcuda_destroy(c); };
  {
    std::vector<std::thread> creators;
    for (int g = 0; g < gpus; ++g)
      creators.emplace_back([&, g]() {
        // shard g runs on the g-th sm_100 device of the box (CUDA ordinals need not be 0..N-1 on a mixed box)
        if (mvgcuda_create(std::max(0, mvgcuda_device_ordinal(g % n_devices)), &ctxs[g]) != MVGCUDA_OK) ctx_error[g] = mvgcuda_last_error(NULL);
      });
    for (auto& t : creators) t.join();
  }
  for (int g = 0; g < gpus; ++g)
    if (!ctx_error[g].empty()) { std::cerr << "GPU " << g << ": " << ctx_error[g] << std::endl; destroy_all(); return EXIT_FAILURE; }
  const Clock::time_point t_ctx = Clock::now();
  std::cout << std::endl << " - PUTATIVE MATCHES - " << n << " images, " << n_pairs << " pairs, " << gpus << " GPU(s)" << std::endl;

  // ---- 2. streaming load: parse -> page-locked staging -> every GPU
  for (int g = 0; g < gpus; ++g)
    if (mvgcuda_stream_begin(ctxs[g], n, rows.data()) != MVGCUDA_OK) {
      std::cerr << "GPU " << g << ": " << mvgcuda_last_error(ctxs[g]) << std::endl;
      destroy_all();
      return EXIT_FAILURE;
    }
  // page-locked staging, carved out of per-thread slabs (one cudaHostAlloc per 64 MB instead of one per image); per image:
  // [rows][128] u8 then [rows][2] float; freed after stream_end
  std::vector<void*> staging;
  std::mutex staging_mtx;
  const size_t kSlabBytes = 64ull << 20;
  std::vector<std::mutex> ctx_mtx(gpus);
  std::atomic<int> next(0), first_bad(n), mismatch(0);
  std::string stream_error;
  std::mutex err_mtx;
  auto load = [&]() {
    uint8_t* slab = NULL;
    size_t slab_left = 0;
    std::vector<char> text;
    for (int i = next++; i < n; i = next++) {
      const std::string base = opt.outdir + "/" + basename_part(names[i]);
      const size_t nrows = (size_t)rows[i];
      uint8_t* d = NULL;
      float* xy = NULL;
      if (nrows) {
        const size_t need = (nrows * (MVGCUDA_DIM + 2 * sizeof(float)) + 255) & ~size_t(255);
        if (need > slab_left) {
          void* p_ = NULL;
          const size_t bytes = std::max(need, kSlabBytes);
          if (mvgcuda_host_alloc(bytes, &p_) != MVGCUDA_OK) { first_bad = std::min(first_bad.load(), i); continue; }
          { std::lock_guard<std::mutex> lk(staging_mtx); staging.push_back(p_); }
          slab = static_cast<uint8_t*>(p_);
          slab_left = bytes;
        }
        d = slab;
        slab += need;
        slab_left -= need;
        xy = reinterpret_cast<float*>(d + nrows * MVGCUDA_DIM);
        std::ifstream f((base + ".desc").c_str(), std::ios::binary);
        f.seekg(hdr_bytes[i]);
        f.read(reinterpret_cast<char*>(d), (std::streamsize)(nrows * MVGCUDA_DIM));
        if ((size_t)f.gcount() != nrows * MVGCUDA_DIM) { first_bad = std::min(first_bad.load(), i); continue; }
      }
      // .feat: x y scale orientation per line (feature.h:117-133); the row count of the pair comes from the features in
      // the reference (matcher_all_in_memory.h:80) -- the streaming layout assumes it equals the .desc count
      size_t got = 0;
      {
        std::ifstream f((base + ".feat").c_str(), std::ios::binary | std::ios::ate);
        if (f.is_open()) {
          const std::streamoff size = f.tellg();
          f.seekg(0);
          text.resize((size_t)std::max<std::streamoff>(size, 0));
          if (size > 0) f.read(text.data(), size);
          if (f.bad() || (size > 0 && f.gcount() != size)) { first_bad = std::min(first_bad.load(), i); continue; }
          got = parse_feat_buffer(text.data(), text.data() + text.size(), xy, nrows);
        }
      }
      if (got != nrows) { ++mismatch; continue; }
      for (int g = 0; g < gpus; ++g) {
        std::lock_guard<std::mutex> lk(ctx_mtx[g]);
        if (mvgcuda_stream_image(ctxs[g], i, d, xy) != MVGCUDA_OK) {
          std::lock_guard<std::mutex> el(err_mtx);
          if (stream_error.empty()) stream_error = std::string("GPU ") + std::to_string(g) + ": " + mvgcuda_last_error(ctxs[g]);
        }
      }
    }
  };
  {
    const int n_threads = (int)std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)std::max(n, 1)));
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(load);
    for (auto& t : pool) t.join();
  }
  auto free_staging = [&]() { for (void* p_ : staging) if (p_) mvgcuda_host_free(p_); };
  bool stream_ok = stream_error.empty() && first_bad.load() >= n && mismatch.load() == 0;
  for (int g = 0; g < gpus && stream_ok; ++g)
    if (mvgcuda_stream_end(ctxs[g]) != MVGCUDA_OK) { stream_error = mvgcuda_last_error(ctxs[g]); stream_ok = false; }
  if (first_bad.load() < n) {
    std::cerr << "cannot parse " << opt.outdir << "/" << basename_part(names[first_bad.load()]) << ".feat/.desc" << std::endl;
    for (int g = 0; g < gpus; ++g) mvgcuda_stream_end(ctxs[g]);
    free_staging(); destroy_all();
    return EXIT_FAILURE;
  }
  if (!stream_error.empty()) { std::cerr << stream_error << std::endl; free_staging(); destroy_all(); return EXIT_FAILURE; }
  if (mismatch.load() > 0) {
    // a .feat with a different row count than its .desc: the reference takes the count from the features and reads past /
    // short of the descriptors; refused here rather than reproduced
    std::cerr << mismatch.load() << " image(s) whose .feat and .desc row counts differ: refused (the reference over-/under-reads here)" << std::endl;
    for (int g = 0; g < gpus; ++g) mvgcuda_stream_end(ctxs[g]);
    free_staging(); destroy_all();
    return EXIT_FAILURE;
  }
  free_staging();
  const Clock::time_point t_load = Clock::now();

  // putative matches of the whole collection for the filter: pairs in std::map order, counts, offsets, (_i, _j)
  std::vector<int32_t> put_pairs, put_counts, put_matches;
  std::vector<int64_t> put_offsets;
  Clock::time_point t_match = t_load, t_export = t_load;
  long long total = 0;
  float gpu_ms_max = 0;
  if (!resumed) {
  // ---- 3. pair shards
  std::vector<int32_t> pairs;
  pairs.reserve((size_t)n_pairs * 2);
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) { pairs.push_back(i); pairs.push_back(j); }
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
  struct Shard { std::vector<int32_t> counts; std::vector<int32_t> matches; std::string error; float gpu_ms = 0; std::vector<std::string> text; long long total = 0; };
  const int fmt_threads = (int)std::max(1u, std::thread::hardware_concurrency() / (unsigned)gpus);  // per shard
  std::vector<Shard> shards(gpus);
  auto work = [&](int g) {
    Shard& S = shards[g];
    const int64_t b = bounds[g], e = bounds[g + 1];
    mvgcuda_pair_matches pm;
    if (mvgcuda_match_collection(ctxs[g], e - b, pairs.data() + 2 * b, ratio_sq, 0, &pm) != MVGCUDA_OK) {
      S.error = mvgcuda_last_error(ctxs[g]);
      return;
    }
    S.gpu_ms = pm.gpu_ms;
    if (filter_f) {  // the filter needs the lists themselves (the library's buffers are reused by the next call)
      S.counts.assign(pm.counts, pm.counts + (e - b));
      S.matches.assign(pm.matches, pm.matches + 2 * pm.offsets[e - b]);
    }
    // the text of the shard, straight from the library's result buffers (pairs are in lexicographic (i, j) order ==
    // std::map iteration order), formatted by fmt_threads threads over chunks of equal match count
    const int64_t np_ = e - b;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(fmt_threads, np_));
    S.text.assign(nt, std::string());
    std::vector<int64_t> cut(nt + 1, np_);
    cut[0] = 0;
    for (int t = 1; t < nt; ++t) {
      const long long target = pm.offsets[np_] * t / nt;
      cut[t] = std::lower_bound(pm.offsets, pm.offsets + np_, target) - pm.offsets;
    }
    auto fmt = [&](int t) {
      std::string& out = S.text[t];
      const int64_t p0 = b + cut[t], p1 = b + cut[t + 1];
      out.reserve((size_t)(pm.offsets[p1 - b] - pm.offsets[p0 - b]) * 12 + (size_t)(p1 - p0) * 16 + 64);
      for (int64_t p = p0; p < p1; ++p) {
        const int c = pm.counts[p - b];
        const int32_t* m = pm.matches + 2 * pm.offsets[p - b];
        append_uint(out, (unsigned)pairs[2 * p]); out.push_back(' ');
        append_uint(out, (unsigned)pairs[2 * p + 1]); out.push_back('\n');
        append_uint(out, (unsigned)c); out.push_back('\n');
        for (int k = 0; k < c; ++k) {
          append_uint(out, (unsigned)m[2 * k]); out.push_back(' ');
          append_uint(out, (unsigned)m[2 * k + 1]); out.push_back('\n');
        }
      }
    };
    {
      std::vector<std::thread> ft;
      for (int t = 1; t < nt; ++t) ft.emplace_back(fmt, t);
      fmt(0);
      for (auto& t : ft) t.join();
    }
    S.total = pm.offsets[np_];
  };
  {
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g) th.emplace_back(work, g);
    for (auto& t : th) t.join();
  }
  t_match = Clock::now();
  for (int g = 0; g < gpus; ++g)
    if (!shards[g].error.empty()) { std::cerr << "GPU " << g << ": " << shards[g].error << std::endl; destroy_all(); return EXIT_FAILURE; }

  // ---- 4. export in shard (== pair) order
  FILE* f = fopen(putative.c_str(), "wb");
  if (!f) { std::cerr << "cannot write " << putative << std::endl; destroy_all(); return EXIT_FAILURE; }
  bool ok = true;
  for (int g = 0; g < gpus; ++g) {
    for (const std::string& chunk : shards[g].text) ok = ok && fwrite(chunk.data(), 1, chunk.size(), f) == chunk.size();
    total += shards[g].total;
  }
  if (fclose(f) != 0 || !ok) { std::cerr << "short write to " << putative << std::endl; destroy_all(); return EXIT_FAILURE; }
  t_export = Clock::now();
  for (int g = 0; g < gpus; ++g) gpu_ms_max = std::max(gpu_ms_max, shards[g].gpu_ms);
  if (filter_f) {
    put_pairs = pairs;
    put_counts.reserve((size_t)n_pairs);
    put_matches.reserve((size_t)total * 2);
    for (int g = 0; g < gpus; ++g) {
      put_counts.insert(put_counts.end(), shards[g].counts.begin(), shards[g].counts.end());
      put_matches.insert(put_matches.end(), shards[g].matches.begin(), shards[g].matches.end());
    }
  }
  } else {
    // imported putatives with the map semantics of pairedIndexedMatchImport: distinct keys in order, the LAST block of a key wins
    std::vector<size_t> order(im.blocks.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return im.blocks[x].first < im.blocks[y].first; });
    for (size_t k = 0; k < order.size(); ++k) {
      if (k + 1 < order.size() && im.blocks[order[k + 1]].first == im.blocks[order[k]].first) continue;
      const auto& blk = im.blocks[order[k]];
      if (blk.first.first >= (size_t)n || blk.first.second >= (size_t)n) { std::cerr << "imported pair out of range" << std::endl; destroy_all(); return EXIT_FAILURE; }
      put_pairs.push_back((int32_t)blk.first.first);
      put_pairs.push_back((int32_t)blk.first.second);
      put_counts.push_back((int32_t)blk.second.size());
      for (const auto& m : blk.second) { put_matches.push_back(m.first); put_matches.push_back(m.second); }
    }
  }
  if (!resumed)
    std::cout << "start-up (contexts + headers) " << ms(t_begin, t_ctx) << " ms; load (parse + upload to " << gpus << " GPU(s)) " << ms(t_ctx, t_load)
              << " ms; match + format " << ms(t_load, t_match) << " ms (GPU time " << gpu_ms_max << " ms, " << n_pairs << " pairs, " << total
              << " putative matches, " << (n_pairs / std::max(1e-9, ms(t_load, t_match) * 1e-3)) << " pairs/s); export " << ms(t_match, t_export)
              << " ms -> " << putative << std::endl;
  if (!filter_f) {
    destroy_all();
    std::cout << "geometric filtering with -g " << opt.geometric_model << " is not part of this build (only -g f and -g h are)" << std::endl;
    return EXIT_SUCCESS;
  }
  // ---- 5. geometric filter (fundamental matrix or homography, AC-RANSAC) on GPU 0
  std::cout << std::endl << " - GEOMETRIC FILTERING - " << std::endl;
  const size_t np = put_counts.size();
  put_offsets.assign(np + 1, 0);
  for (size_t k = 0; k < np; ++k) put_offsets[k + 1] = put_offsets[k] + put_counts[k];
  if (put_matches.empty()) put_matches.assign(2, 0);
  mvgcuda_pair_matches gm;
  const double max_residual_error = 4.0;  // compute_matches.cpp:254
  if (mvgcuda_geometric_filter(ctxs[0], geo_model, max_residual_error, 4096, 1u, (int64_t)np, put_pairs.data(), put_counts.data(), put_offsets.data(),
                               put_matches.data(), image_sizes.data(), &gm) != MVGCUDA_OK) {
    std::cerr << "geometric filter: " << mvgcuda_last_error(ctxs[0]) << std::endl;
    destroy_all();
    return EXIT_FAILURE;
  }
  const Clock::time_point t_geo = Clock::now();
  const std::string geo_path = opt.outdir + (geo_model == 'h' ? "/matches.h.txt" : "/matches.f.txt");  // compute_matches.cpp:99-113
  if (mvgcuda_write_matches(geo_path.c_str(), gm.n_pairs, put_pairs.data(), gm.counts, gm.offsets, gm.matches, 1) != MVGCUDA_OK) {
    std::cerr << "cannot write " << geo_path << std::endl;
    destroy_all();
    return EXIT_FAILURE;
  }
  long long kept_pairs = 0;
  for (int64_t k = 0; k < gm.n_pairs; ++k) kept_pairs += gm.counts[k] > 0;
  std::cout << "geometric filter (" << (geo_model == 'h' ? "H" : "F") << ", AC-RANSAC, 4096 iterations, " << max_residual_error << " px) " << ms(t_export, t_geo) << " ms (GPU time " << gm.gpu_ms << " ms): "
            << kept_pairs << " of " << np << " pairs kept, " << gm.offsets[gm.n_pairs] << " matches, " << gm.rescanned_queries
            << " rand() values consumed, " << gm.knn_kernel_launches << " models re-evaluated with host roots -> " << geo_path << std::endl;
  destroy_all();
  return EXIT_SUCCESS;
}
