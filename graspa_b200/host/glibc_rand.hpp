// graspa_b200 host layer -- the reference's random stream without the libc lock.
//
// The reference draws every uniform from glibc's rand() (data_struct.cpp:6-11: rand()/RAND_MAX, seeded by
// std::srand(RANDOMSEED), data_struct.h:1340).  glibc's default generator is the TYPE_3 additive feedback generator
// r[i] = r[i-3] + r[i-31] (mod 2^32), output r[i] >> 1.  Restating it here makes the stream (a) independent of the
// C library the host program is linked against, (b) peekable -- the fused move calls pass uniforms that are consumed
// only when the stage they belong to survives -- and (c) ~10x faster than locked rand() calls when a pool refill
// draws 1 000 002 values.  tests/test_host_driver.py pins it against the reference's stream (tests/golden/rng_seed0.npz).
#pragma once
#include <cstdint>
#include <vector>

class GlibcRand
{
 public:
  explicit GlibcRand(unsigned seed = 1) { reseed(seed); }
  void reseed(unsigned seed)
  {
    if(seed == 0) seed = 1;                         // srandom_r: "We must make sure the seed is not 0"
    int32_t r[34];
    r[0] = (int32_t) seed;
    for(int i = 1; i < 31; i++)
    {
      // 16807 * r[i-1] % 2147483647 without overflow (Schrage), as glibc does
      const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
      long word = 16807 * lo - 2836 * hi;
      if(word < 0) word += 2147483647;
      r[i] = (int32_t) word;
    }
    for(int i = 0; i < 31; i++) s_[i] = (uint32_t) r[i];
    f_ = 3; b_ = 0;
    for(int i = 0; i < 310; i++) next_raw();        // srandom_r discards 10 * rand_deg outputs
    buf_.clear(); pos_ = 0;
  }
  // the reference's Get_Uniform_Random()
  double uniform() { return (double) next() / 2147483647.0; }
  // value the k-th next call to uniform() will return, without consuming anything
  double peek(size_t k = 0)
  {
    while(buf_.size() - pos_ <= k) buf_.push_back(next_raw());
    return (double) buf_[pos_ + k] / 2147483647.0;
  }
  void advance(size_t n) { for(size_t i = 0; i < n; i++) next(); }
  uint64_t consumed() const { return count_; }

 private:
  uint32_t next_raw()
  {
    s_[f_] += s_[b_];
    const uint32_t out = s_[f_] >> 1;
    if(++f_ == 31) f_ = 0;
    if(++b_ == 31) b_ = 0;
    return out;
  }
  uint32_t next()
  {
    count_++;
    if(pos_ < buf_.size())
    {
      const uint32_t v = buf_[pos_++];
      if(pos_ == buf_.size()) { buf_.clear(); pos_ = 0; }
      return v;
    }
    return next_raw();
  }
  uint32_t s_[31];
  int f_ = 3, b_ = 0;
  std::vector<uint32_t> buf_; size_t pos_ = 0;
  uint64_t count_ = 0;
};
