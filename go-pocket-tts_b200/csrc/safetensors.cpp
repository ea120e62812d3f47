// safetensors.cpp — minimal .safetensors reader for non-Go hosts of the C ABI
// (the Go backend keeps using internal/safetensors and calls ptts_load_tensor
// per key).  Format per reference internal/safetensors/store.go:246-271:
// 8-byte little-endian header length, JSON header {name: {dtype, shape,
// data_offsets}, "__metadata__": {...}}, then the raw tensor bytes.
#include "engine.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace ptts {
namespace {

struct Cur {
  const char *p, *end;
  bool ok = true;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r'))
      p++;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      p++;
      return true;
    }
    return false;
  }
  std::string str() {
    ws();
    std::string s;
    if (p >= end || *p != '"') {
      ok = false;
      return s;
    }
    p++;
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        p++;
        switch (*p) {
        case 'n': s += '\n'; break;
        case 't': s += '\t'; break;
        case 'u': // keep escaped code points opaque; tensor names are ASCII
          s += "\\u";
          break;
        default: s += *p;
        }
        p++;
      } else
        s += *p++;
    }
    if (p >= end) {
      ok = false;
      return s;
    }
    p++;
    return s;
  }
  long long num() {
    ws();
    char *q = nullptr;
    long long v = strtoll(p, &q, 10);
    if (q == p)
      ok = false;
    p = q;
    return v;
  }
  void skip_value() { // strings, numbers, flat/nested objects and arrays
    ws();
    if (p >= end) {
      ok = false;
      return;
    }
    if (*p == '"') {
      str();
    } else if (*p == '{' || *p == '[') {
      char open = *p, close = open == '{' ? '}' : ']';
      p++;
      ws();
      if (eat(close))
        return;
      do {
        if (open == '{') {
          str();
          if (!eat(':')) {
            ok = false;
            return;
          }
        }
        skip_value();
        if (!ok)
          return;
      } while (eat(','));
      if (!eat(close))
        ok = false;
    } else {
      while (p < end && *p != ',' && *p != '}' && *p != ']')
        p++;
    }
  }
};

} // namespace

int load_safetensors_file(Engine *e, const char *path) {
  int fd = open(path, O_RDONLY);
  if (fd < 0)
    PTTS_FAIL(PTTS_EINVAL, "safetensors: open %s failed", path);
  struct stat sb;
  if (fstat(fd, &sb) != 0 || sb.st_size < 8) {
    close(fd);
    PTTS_FAIL(PTTS_EINVAL, "safetensors: file too small");
  }
  size_t size = (size_t)sb.st_size;
  void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED)
    PTTS_FAIL(PTTS_EINVAL, "safetensors: mmap failed");
  const unsigned char *base = (const unsigned char *)map;
  uint64_t hl = 0;
  for (int i = 7; i >= 0; i--)
    hl = (hl << 8) | base[i];
  if (hl == 0 || 8 + hl > size) {
    munmap(map, size);
    PTTS_FAIL(PTTS_EINVAL, "safetensors: invalid header length %llu", (unsigned long long)hl);
  }
  const unsigned char *data = base + 8 + hl;
  size_t data_size = size - 8 - (size_t)hl;
  Cur c{(const char *)base + 8, (const char *)base + 8 + hl};
  int rc = PTTS_OK, count = 0;
  if (!c.eat('{')) {
    munmap(map, size);
    PTTS_FAIL(PTTS_EINVAL, "safetensors: parse header: expected object");
  }
  c.ws();
  if (!c.eat('}')) {
    do {
      std::string name = c.str();
      if (!c.ok || !c.eat(':')) {
        rc = PTTS_EINVAL;
        set_error("safetensors: parse header: malformed entry");
        break;
      }
      if (name == "__metadata__") {
        c.skip_value();
        continue;
      }
      if (!c.eat('{')) {
        rc = PTTS_EINVAL;
        set_error("safetensors: parse header: tensor %s is not an object", name.c_str());
        break;
      }
      std::string dtype;
      std::vector<int64_t> shape;
      long long o0 = -1, o1 = -1;
      do {
        std::string k = c.str();
        if (!c.eat(':')) {
          c.ok = false;
          break;
        }
        if (k == "dtype") {
          dtype = c.str();
        } else if (k == "shape") {
          if (!c.eat('[')) {
            c.ok = false;
            break;
          }
          c.ws();
          if (!c.eat(']')) {
            do {
              shape.push_back(c.num());
            } while (c.eat(','));
            if (!c.eat(']'))
              c.ok = false;
          }
        } else if (k == "data_offsets") {
          if (!c.eat('[')) {
            c.ok = false;
            break;
          }
          o0 = c.num();
          if (!c.eat(','))
            c.ok = false;
          o1 = c.num();
          if (!c.eat(']'))
            c.ok = false;
        } else {
          c.skip_value();
        }
      } while (c.ok && c.eat(','));
      if (!c.ok || !c.eat('}')) {
        rc = PTTS_EINVAL;
        set_error("safetensors: parse header: malformed tensor %s", name.c_str());
        break;
      }
      int dt;
      size_t esz;
      if (dtype == "F32") {
        dt = PTTS_DTYPE_F32;
        esz = 4;
      } else if (dtype == "F16") {
        dt = PTTS_DTYPE_F16;
        esz = 2;
      } else if (dtype == "BF16") {
        dt = PTTS_DTYPE_BF16;
        esz = 2;
      } else {
        continue; // e.g. I64 bookkeeping tensors are not model weights
      }
      size_t n = 1;
      for (int64_t d : shape)
        n *= (size_t)d;
      if (o0 < 0 || o1 < o0 || (size_t)o1 > data_size || (size_t)(o1 - o0) != n * esz) {
        rc = PTTS_EINVAL;
        set_error("safetensors: tensor %s has invalid data_offsets", name.c_str());
        break;
      }
      rc = e->load_tensor(name.c_str(), shape.data(), (int)shape.size(), dt, data + o0);
      if (rc)
        break;
      count++;
    } while (c.eat(','));
  }
  munmap(map, size);
  if (rc == PTTS_OK && count == 0)
    PTTS_FAIL(PTTS_EINVAL, "safetensors: no tensors found");
  return rc;
}

} // namespace ptts
