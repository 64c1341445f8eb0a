// line_pipeline.h -- ordered parallel map over the lines of a (gzip or plain) text file.
//
// The reference walks its genotype text files three times (QC ReadFile_geno src/gemma_io.cpp:639-873, BimbamKin
// :1418-1597, LMM::Analyze src/lmm.cpp:1474-1658), each time one strtok/atof per genotype on a single thread; at
// n = 10^4 x p = 10^6 that is ~10^10 tokens per pass and dwarfs the GPU time.  Here one thread inflates and cuts the
// file into blocks of lines, a pool of workers tokenises / converts the blocks, and the consumer receives the finished
// blocks in file order (so every downstream result keeps the reference's SNP order).
#pragma once
#include <zlib.h>

#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct LineBlock {
  size_t first_line = 0;              // 0-based index of lines[0] among the file's non-blank lines
  std::unique_ptr<char[]> data;       // the block's text; every line NUL-terminated in place
  std::vector<char *> lines;          // line starts inside data (without the trailing "\n" / "\r\n"); workers may write into them
};

inline int host_threads() {
  if (const char *e = std::getenv("GB200_HOST_THREADS")) { const int v = std::atoi(e); if (v > 0) return v; }
  unsigned hc = std::thread::hardware_concurrency();
  if (hc == 0) hc = 4;
  return (int)(hc > 32 ? 32 : hc);
}

template <class Out>
class LinePipeline {
 public:
  using Work = std::function<void(LineBlock &, Out &)>;

  LinePipeline(const std::string &path, Work work, int threads = 0) : work_(std::move(work)) {
    f_ = gzopen(path.c_str(), "rb");
    if (!f_) return;
    gzbuffer(f_, 1 << 20);
    const int nt = threads > 0 ? threads : host_threads();
    cap_ = (size_t)nt * 2 + 2;
    reader_ = std::thread([this] { read_loop(); });
    for (int i = 0; i < nt; ++i) workers_.emplace_back([this] { work_loop(); });
  }
  ~LinePipeline() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_in_.notify_all(); cv_out_.notify_all(); cv_room_.notify_all();
    if (reader_.joinable()) reader_.join();
    for (auto &w : workers_) if (w.joinable()) w.join();
    if (f_) gzclose(f_);
  }
  bool ok() const { return f_ != nullptr; }

  // next finished block, in file order; false at end of file
  bool next(Out &out) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_out_.wait(lk, [&] { return done_.count(consumed_) || (eof_ && consumed_ == produced_) || stop_; });
    auto it = done_.find(consumed_);
    if (it == done_.end()) return false;
    out = std::move(it->second);
    done_.erase(it);
    ++consumed_;
    lk.unlock();
    cv_room_.notify_one();
    return true;
  }

 private:
  void read_loop() {
    // bulk gzread; a block = all complete lines of one 1 MiB read (plus the tail carried over from the previous one)
    std::vector<char> buf(1 << 20);
    size_t line_no = 0;
    std::string carry;
    auto push = [&](const char *a, size_t na, const char *b2, size_t nb) -> bool {      // text = a[0..na) + b2[0..nb), ends with '\n' or EOF
      LineBlock blk; blk.first_line = line_no;
      const size_t len = na + nb;
      blk.data.reset(new char[len + 1]);
      char *d = blk.data.get();
      if (na) std::memcpy(d, a, na);
      if (nb) std::memcpy(d + na, b2, nb);
      d[len] = 0;
      char *p = d, *end = d + len;
      while (p < end) {
        char *nl = (char *)std::memchr(p, '\n', (size_t)(end - p));
        char *stop = nl ? nl : end;
        *stop = 0;
        if (stop > p && stop[-1] == '\r') stop[-1] = 0;
        // lines without a token (empty or delimiters only) are dropped HERE, so that every pass over the file numbers the
        // remaining lines identically (QC records one entry per line it sees; the later passes index by line number)
        const char *q = p;
        while (*q == ' ' || *q == ',' || *q == '\t') ++q;
        if (*q) blk.lines.push_back(p);
        p = stop + 1;
      }
      line_no += blk.lines.size();
      if (blk.lines.empty()) return true;
      std::unique_lock<std::mutex> lk(mu_);
      cv_room_.wait(lk, [&] { return produced_ - consumed_ < cap_ || stop_; });
      if (stop_) return false;
      todo_.emplace_back(produced_++, std::move(blk));
      lk.unlock();
      cv_in_.notify_one();
      return true;
    };
    for (;;) {
      const int got = gzread(f_, buf.data(), (unsigned)buf.size());
      if (got <= 0) break;
      const char *b0 = buf.data();
      const char *last = nullptr;
      for (const char *q = b0 + got; q > b0; --q) if (q[-1] == '\n') { last = q - 1; break; }
      if (!last) { carry.append(b0, (size_t)got); continue; }
      if (!push(carry.data(), carry.size(), b0, (size_t)(last - b0 + 1))) return;
      carry.assign(last + 1, (size_t)(b0 + got - (last + 1)));
    }
    if (!carry.empty() && !push(carry.data(), carry.size(), nullptr, 0)) return;          // last line without a newline
    {
      std::lock_guard<std::mutex> g(mu_);
      eof_ = true;
    }
    cv_in_.notify_all(); cv_out_.notify_all();
  }
  void work_loop() {
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_in_.wait(lk, [&] { return !todo_.empty() || eof_ || stop_; });
      if (stop_) return;
      if (todo_.empty()) { if (eof_) return; continue; }
      auto item = std::move(todo_.front());
      todo_.pop_front();
      lk.unlock();
      Out out;
      work_(item.second, out);
      lk.lock();
      done_.emplace(item.first, std::move(out));
      lk.unlock();
      cv_out_.notify_all();
    }
  }

  Work work_;
  gzFile f_ = nullptr;
  size_t cap_ = 8;
  std::thread reader_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_in_, cv_out_, cv_room_;
  std::deque<std::pair<size_t, LineBlock>> todo_;
  std::map<size_t, Out> done_;
  size_t produced_ = 0, consumed_ = 0;
  bool eof_ = false, stop_ = false;
};

// ---- tokens -------------------------------------------------------------------------------------------------
// The reference tokenises with strtok(line, " ,\t") (src/gemma_io.cpp:706 ff).  next_token() is the re-entrant equivalent:
// it skips leading delimiters, NUL-terminates the token in place and advances the cursor.
inline char *next_token(char *&cur) {
  char *p = cur;
  while (*p == ' ' || *p == ',' || *p == '\t') ++p;
  if (!*p) { cur = p; return nullptr; }
  char *b = p;
  while (*p && *p != ' ' && *p != ',' && *p != '\t') ++p;
  if (*p) { *p = 0; cur = p + 1; } else cur = p;
  return b;
}

// == atof(tok) for every input.  Fast path: [sign] digits [. digits] with at most 15 significant digits and at most 22
// fraction digits: mantissa and power of ten are both exact doubles, so one IEEE division is correctly rounded -- the same
// value strtod returns.  Anything else (exponents, hex, inf/nan, garbage) goes to strtod.
inline double token_to_double(const char *tok) {
  static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char *p = tok;
  bool neg = false;
  if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
  uint64_t m = 0; int sig = 0, frac = 0; bool any = false;
  while (*p >= '0' && *p <= '9') { any = true; if (m || *p != '0') ++sig; m = m * 10 + (uint64_t)(*p - '0'); if (sig > 15) return std::atof(tok); ++p; }
  if (*p == '.') {
    ++p;
    while (*p >= '0' && *p <= '9') { any = true; if (m || *p != '0') ++sig; m = m * 10 + (uint64_t)(*p - '0'); ++frac; if (sig > 15 || frac > 22) return std::atof(tok); ++p; }
  }
  if (*p != 0 || !any) return std::atof(tok);
  const double d = (double)m / p10[frac];
  return neg ? -d : d;
}
