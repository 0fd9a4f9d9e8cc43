// b200kv_remote.cpp — remote chunk tier: the stand-in for LMCache's cache server
// (`/opt/venv/bin/lmcache_server 0.0.0.0 <port>`, helm/templates/deployment-cache-server.yaml:62-65)
// and for the `LMCACHE_REMOTE_URL=lm://host:port` client side the engines are configured with
// (helm/templates/deployment-vllm-multi.yaml:338-345; helm/tests/lmcache_test.yaml:166-181).
//
// CPU only.  The server keeps chunks in an ordinary b200kv pool (same index,
// LRU and pinning as the engines' host pool); the client moves a chunk between a socket and a slot
// of the LOCAL pinned pool with no intermediate buffer, so a fetched chunk is immediately loadable by
// the GPU path.  The wire format is this library's own (LMCache's is not in /root/reference —
// parity unpinned, like the rest of the KV path): one fixed 48-byte little-endian header per frame,
// then `length` payload bytes.
//
//   PING   -> status
//   EXISTS (payload: n keys)            -> n_tokens field = length of the present PREFIX of the keys
//   GET    key                          -> status, fmt, n_tokens, length, payload
//   PUT    key fmt n_tokens length      -> status (0 = send it, -EEXIST = already there) ; payload -> status
//   STATS                               -> payload: b200kv_pool_stats
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "b200kv.h"

namespace {

constexpr uint32_t kMagic = 0x564b3242u;  // "B2KV"
constexpr uint16_t kVersion = 1;
enum Op : uint16_t { kPing = 1, kPut = 2, kGet = 3, kExists = 4, kStats = 5 };
constexpr uint64_t kMaxKeysPerExists = 1u << 16;

struct Frame {
  uint32_t magic;
  uint16_t version;
  uint16_t op;
  uint64_t key;
  uint32_t fmt;
  int32_t n_tokens;
  uint32_t owner;
  int32_t status;
  uint64_t length;      // payload bytes that follow this header
  uint64_t slot_bytes;  // sender's chunk slot size (PUT) / server's (GET reply)
};
static_assert(sizeof(Frame) == 48, "wire header is 48 bytes");

Frame make_frame(uint16_t op) {
  Frame f;
  memset(&f, 0, sizeof(f));
  f.magic = kMagic;
  f.version = kVersion;
  f.op = op;
  return f;
}

bool send_all(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n) {
    const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}

bool recv_all(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n) {
    const ssize_t k = ::recv(fd, p, n, 0);
    if (k == 0) return false;
    if (k < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}

bool drain(int fd, uint64_t n) {
  std::vector<char> scratch(1 << 20);
  while (n) {
    const size_t k = static_cast<size_t>(n < scratch.size() ? n : scratch.size());
    if (!recv_all(fd, scratch.data(), k)) return false;
    n -= k;
  }
  return true;
}

bool recv_frame(int fd, Frame* f) {
  return recv_all(fd, f, sizeof(*f)) && f->magic == kMagic && f->version == kVersion;
}

void tune_socket(int fd, int timeout_ms) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  // B200KV_SOCKBUF_KB > 0 pins the socket buffers (long fat links); default: leave the kernel's
  // autotuning on — a fixed size is clamped to net.core.{w,r}mem_max and switches autotuning off.
  const char* e = getenv("B200KV_SOCKBUF_KB");
  int buf = e ? atoi(e) * 1024 : 0;
  if (buf > 0) {
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
  }
  if (timeout_ms > 0) {
    timeval tv;
    tv.tv_sec = timeout_ms / 1000;
    tv.tv_usec = (timeout_ms % 1000) * 1000;
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
  }
}

uint64_t slot_bytes_of(b200kv_pool* pool) {
  b200kv_pool_stats st;
  if (b200kv_pool_get_stats(pool, &st) != B200KV_OK) return 0;
  return st.slot_bytes;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// server
// ------------------------------------------------------------------------------------------------
struct b200kv_server {
  int listen_fd = -1;
  int wake_pipe[2] = {-1, -1};
  int port = 0;
  uint64_t pool_bytes = 0;
  std::mutex mu;               // guards pool creation, conns
  // One pool per chunk geometry, created by the first PUT of that slot size: the chart runs ONE cache
  // server for every model it deploys (modelSpec is a list), and chunk sizes differ between models.
  std::vector<std::pair<uint64_t, b200kv_pool*>> pools;
  std::set<int> conns;
  std::atomic<bool> stopping{false};
  std::atomic<int> live_threads{0};
  std::thread acceptor;
  std::atomic<uint64_t> n_put{0}, n_get{0}, n_get_miss{0}, bytes_in{0}, bytes_out{0};

  // the pool of this slot size (created on demand), or nullptr
  b200kv_pool* pool_for(uint64_t slot_bytes) {
    if (!slot_bytes || slot_bytes % 16) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    for (auto& p : pools)
      if (p.first == slot_bytes) return p.second;
    // B200KV_SERVER_GB is the budget of the whole server, not of each geometry: it is split evenly over the
    // B200KV_SERVER_GEOMETRIES (default 1) chunk geometries the deployment serves, and a geometry beyond
    // that number is refused (its PUTs fail, the engines treat the server as a miss) — a client cannot
    // make the server allocate a pool per slot size it invents.
    const char* ge = getenv("B200KV_SERVER_GEOMETRIES");
    const size_t max_geoms = ge && atoi(ge) > 0 ? static_cast<size_t>(atoi(ge)) : 1;
    if (pools.size() >= max_geoms || pools.size() >= 16) return nullptr;
    const uint64_t share = pool_bytes / max_geoms;
    b200kv_pool_config pc;
    memset(&pc, 0, sizeof(pc));
    pc.shm_name = nullptr;
    pc.pool_bytes = share < slot_bytes ? slot_bytes : share;
    pc.slot_bytes = slot_bytes;
    pc.flags = B200KV_POOL_CREATE;
    b200kv_pool* np = nullptr;
    if (b200kv_pool_open(&pc, &np) != B200KV_OK) return nullptr;
    pools.emplace_back(slot_bytes, np);
    return np;
  }
  std::vector<b200kv_pool*> all_pools() {
    std::lock_guard<std::mutex> lk(mu);
    std::vector<b200kv_pool*> v;
    for (auto& p : pools) v.push_back(p.second);
    return v;
  }
  // keys are namespaced by model and format, so a key lives in at most one pool
  b200kv_pool* pool_holding(uint64_t key) {
    for (b200kv_pool* p : all_pools()) {
      int32_t hit = 0;
      uint32_t owner = 0;
      if (b200kv_pool_lookup_owner(p, &key, 1, &hit, &owner) == B200KV_OK && hit == 1) return p;
    }
    return nullptr;
  }

  void serve(int fd) {
    Frame in;
    while (!stopping.load() && recv_frame(fd, &in)) {
      Frame out = make_frame(in.op);
      out.key = in.key;
      switch (in.op) {
        case kPing:
          if (in.length && !drain(fd, in.length)) return;
          if (!send_all(fd, &out, sizeof(out))) return;
          break;
        case kExists: {
          if (in.length % 8 || in.length / 8 > kMaxKeysPerExists) return;
          std::vector<uint64_t> keys(in.length / 8);
          if (in.length && !recv_all(fd, keys.data(), in.length)) return;
          int32_t hit = 0;
          b200kv_pool* p = keys.empty() ? nullptr : pool_holding(keys[0]);
          if (p) {
            std::vector<uint32_t> owners(keys.size());
            b200kv_pool_lookup_owner(p, keys.data(), static_cast<int32_t>(keys.size()), &hit, owners.data());
          }
          out.n_tokens = hit;
          if (!send_all(fd, &out, sizeof(out))) return;
          break;
        }
        case kGet: {
          if (in.length && !drain(fd, in.length)) return;
          b200kv_pool* p = pool_holding(in.key);
          uint32_t slot = 0, fmt = 0;
          int32_t n_tok = 0;
          if (!p || b200kv_pool_acquire(p, in.key, &slot, &n_tok, &fmt) != B200KV_OK) {
            out.status = B200KV_ENOENT;
            ++n_get_miss;
            if (!send_all(fd, &out, sizeof(out))) return;
            break;
          }
          out.fmt = fmt;
          out.n_tokens = n_tok;
          out.slot_bytes = out.length = slot_bytes_of(p);
          const bool ok = send_all(fd, &out, sizeof(out)) && send_all(fd, b200kv_pool_slot_ptr(p, slot), out.length);
          b200kv_pool_release(p, in.key);
          if (!ok) return;
          ++n_get;
          bytes_out += out.length;
          break;
        }
        case kPut: {
          b200kv_pool* p = pool_for(in.slot_bytes);
          uint32_t slot = 0;
          int rc = B200KV_EINVAL;
          if (p && in.length == in.slot_bytes && in.n_tokens > 0)
            rc = b200kv_pool_reserve(p, in.key, in.n_tokens, in.fmt, in.owner, &slot);
          out.status = rc;
          if (!send_all(fd, &out, sizeof(out))) {
            if (rc == B200KV_OK) b200kv_pool_abort(p, in.key);
            return;
          }
          if (rc != B200KV_OK) break;  // the client does not send the payload
          if (!recv_all(fd, b200kv_pool_slot_ptr(p, slot), in.length)) {
            b200kv_pool_abort(p, in.key);
            return;
          }
          out.status = b200kv_pool_commit(p, in.key);
          ++n_put;
          bytes_in += in.length;
          if (!send_all(fd, &out, sizeof(out))) return;
          break;
        }
        case kStats: {
          if (in.length && !drain(fd, in.length)) return;
          b200kv_pool_stats st;   // summed over the pools; slot_bytes = the requester's geometry, if given
          memset(&st, 0, sizeof(st));
          for (b200kv_pool* p : all_pools()) {
            b200kv_pool_stats one;
            if (b200kv_pool_get_stats(p, &one) != B200KV_OK) continue;
            if (in.slot_bytes && one.slot_bytes != in.slot_bytes) continue;
            st.n_slots += one.n_slots;
            st.n_used += one.n_used;
            st.slot_bytes = one.slot_bytes;
            st.n_lookups += one.n_lookups;
            st.n_stored_chunks += one.n_stored_chunks;
            st.n_evicted_chunks += one.n_evicted_chunks;
            st.n_dropped_chunks += one.n_dropped_chunks;
          }
          out.length = sizeof(st);
          if (!send_all(fd, &out, sizeof(out)) || !send_all(fd, &st, sizeof(st))) return;
          break;
        }
        default:
          return;  // unknown op: drop the connection
      }
    }
  }

  void accept_loop() {
    while (!stopping.load()) {
      pollfd pf[2] = {{listen_fd, POLLIN, 0}, {wake_pipe[0], POLLIN, 0}};
      if (::poll(pf, 2, -1) < 0) {
        if (errno == EINTR) continue;
        break;
      }
      if (pf[1].revents) break;
      if (!(pf[0].revents & POLLIN)) continue;
      const int fd = ::accept(listen_fd, nullptr, nullptr);
      if (fd < 0) continue;
      tune_socket(fd, 0);
      {
        std::lock_guard<std::mutex> lk(mu);
        conns.insert(fd);
      }
      ++live_threads;
      std::thread([this, fd] {
        serve(fd);
        {
          std::lock_guard<std::mutex> lk(mu);
          conns.erase(fd);
        }
        ::close(fd);
        --live_threads;
      }).detach();
    }
  }
};

extern "C" int b200kv_server_start(const char* host, int port, uint64_t pool_bytes, b200kv_server** out) {
  if (!out || port < 0 || port > 65535) return B200KV_EINVAL;
  sockaddr_in addr;
  memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_port = htons(static_cast<uint16_t>(port));
  if (!host || !*host || !strcmp(host, "0.0.0.0")) {
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
  } else if (inet_pton(AF_INET, host, &addr.sin_addr) != 1) {
    return B200KV_EINVAL;
  }
  const int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) return -errno;
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  if (::bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) < 0 || ::listen(fd, 128) < 0) {
    const int e = errno;
    ::close(fd);
    return -e;
  }
  socklen_t len = sizeof(addr);
  getsockname(fd, reinterpret_cast<sockaddr*>(&addr), &len);
  b200kv_server* s = new (std::nothrow) b200kv_server();
  if (!s) {
    ::close(fd);
    return B200KV_ENOMEM;
  }
  if (::pipe2(s->wake_pipe, O_CLOEXEC) < 0) {
    const int e = errno;
    ::close(fd);
    delete s;
    return -e;
  }
  s->listen_fd = fd;
  s->port = ntohs(addr.sin_port);
  s->pool_bytes = pool_bytes;
  s->acceptor = std::thread([s] { s->accept_loop(); });
  *out = s;
  return B200KV_OK;
}

extern "C" int b200kv_server_port(b200kv_server* s) { return s ? s->port : B200KV_EINVAL; }

extern "C" int b200kv_server_get_stats(b200kv_server* s, uint64_t* out5) {
  if (!s || !out5) return B200KV_EINVAL;
  out5[0] = s->n_put.load();
  out5[1] = s->n_get.load();
  out5[2] = s->n_get_miss.load();
  out5[3] = s->bytes_in.load();
  out5[4] = s->bytes_out.load();
  return B200KV_OK;
}

extern "C" int b200kv_server_stop(b200kv_server* s) {
  if (!s) return B200KV_EINVAL;
  s->stopping.store(true);
  const char b = 1;
  if (::write(s->wake_pipe[1], &b, 1) < 0) { /* acceptor also wakes on close */ }
  if (s->acceptor.joinable()) s->acceptor.join();
  {
    std::lock_guard<std::mutex> lk(s->mu);
    for (int fd : s->conns) ::shutdown(fd, SHUT_RDWR);  // unblocks the connection threads
  }
  for (int i = 0; i < 5000 && s->live_threads.load() > 0; ++i) ::usleep(1000);
  ::close(s->listen_fd);
  ::close(s->wake_pipe[0]);
  ::close(s->wake_pipe[1]);
  if (s->live_threads.load() == 0) {
    for (auto& p : s->pools) b200kv_pool_close(p.second);
    delete s;
  }  // else: a connection thread is stuck in the kernel; leak the object rather than free it under the thread
  return B200KV_OK;
}

// ------------------------------------------------------------------------------------------------
// client
// ------------------------------------------------------------------------------------------------
struct b200kv_remote {
  int fd = -1;
  std::mutex mu;  // one request at a time per connection
  uint64_t bytes_up = 0, bytes_down = 0;
};

extern "C" int b200kv_remote_connect(const char* host, int port, int timeout_ms, b200kv_remote** out) {
  if (!host || !out || port <= 0 || port > 65535) return B200KV_EINVAL;
  addrinfo hints;
  memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  addrinfo* res = nullptr;
  const std::string ps = std::to_string(port);
  if (getaddrinfo(host, ps.c_str(), &hints, &res) != 0 || !res) return B200KV_ENOENT;
  int fd = -1, err = ECONNREFUSED;
  for (addrinfo* ai = res; ai; ai = ai->ai_next) {
    fd = ::socket(ai->ai_family, ai->ai_socktype | SOCK_CLOEXEC, ai->ai_protocol);
    if (fd < 0) continue;
    tune_socket(fd, timeout_ms);  // SO_SNDTIMEO also bounds connect()
    if (::connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) break;
    err = errno;
    ::close(fd);
    fd = -1;
  }
  freeaddrinfo(res);
  if (fd < 0) return -err;
  b200kv_remote* r = new (std::nothrow) b200kv_remote();
  if (!r) {
    ::close(fd);
    return B200KV_ENOMEM;
  }
  r->fd = fd;
  *out = r;
  return B200KV_OK;
}

extern "C" int b200kv_remote_close(b200kv_remote* r) {
  if (!r) return B200KV_EINVAL;
  if (r->fd >= 0) ::close(r->fd);
  delete r;
  return B200KV_OK;
}

extern "C" int b200kv_remote_ping(b200kv_remote* r) {
  if (!r) return B200KV_EINVAL;
  std::lock_guard<std::mutex> lk(r->mu);
  Frame f = make_frame(kPing), rep;
  if (!send_all(r->fd, &f, sizeof(f)) || !recv_frame(r->fd, &rep)) return B200KV_EIO;
  return rep.status;
}

extern "C" int b200kv_remote_exists(b200kv_remote* r, const uint64_t* keys, int32_t n_keys, int32_t* n_prefix) {
  if (!r || !n_prefix || n_keys < 0 || (n_keys && !keys) || static_cast<uint64_t>(n_keys) > kMaxKeysPerExists)
    return B200KV_EINVAL;
  *n_prefix = 0;
  if (n_keys == 0) return B200KV_OK;
  std::lock_guard<std::mutex> lk(r->mu);
  Frame f = make_frame(kExists), rep;
  f.length = static_cast<uint64_t>(n_keys) * 8;
  if (!send_all(r->fd, &f, sizeof(f)) || !send_all(r->fd, keys, f.length) || !recv_frame(r->fd, &rep))
    return B200KV_EIO;
  if (rep.status != B200KV_OK) return rep.status;
  *n_prefix = rep.n_tokens;
  return B200KV_OK;
}

// Local pool -> server.  -ENOENT: not (yet) READY locally; -EEXIST: the server already has it
// (nothing is sent); -ENOSPC: the server could not make room.
extern "C" int b200kv_remote_put(b200kv_remote* r, b200kv_pool* local, uint64_t key, uint32_t owner) {
  if (!r || !local) return B200KV_EINVAL;
  uint32_t slot = 0, fmt = 0;
  int32_t n_tok = 0;
  int rc = b200kv_pool_acquire(local, key, &slot, &n_tok, &fmt);
  if (rc != B200KV_OK) return rc;
  const uint64_t sb = slot_bytes_of(local);
  {
    std::lock_guard<std::mutex> lk(r->mu);
    Frame f = make_frame(kPut), rep;
    f.key = key;
    f.fmt = fmt;
    f.n_tokens = n_tok;
    f.owner = owner;
    f.length = f.slot_bytes = sb;
    if (!send_all(r->fd, &f, sizeof(f)) || !recv_frame(r->fd, &rep)) {
      rc = B200KV_EIO;
    } else if (rep.status != B200KV_OK) {
      rc = rep.status;
    } else if (!send_all(r->fd, b200kv_pool_slot_ptr(local, slot), sb) || !recv_frame(r->fd, &rep)) {
      rc = B200KV_EIO;
    } else {
      rc = rep.status;
      if (rc == B200KV_OK) r->bytes_up += sb;
    }
  }
  b200kv_pool_release(local, key);
  return rc;
}

// Server -> local pool (reserve, receive straight into the slot, commit).  OK also when the chunk
// is already local.  -ENOENT: the server does not have it; -ENOSPC: no local slot could be freed.
extern "C" int b200kv_remote_get(b200kv_remote* r, b200kv_pool* local, uint64_t key, uint32_t owner) {
  if (!r || !local) return B200KV_EINVAL;
  const uint64_t sb = slot_bytes_of(local);
  std::lock_guard<std::mutex> lk(r->mu);
  Frame f = make_frame(kGet), rep;
  f.key = key;
  if (!send_all(r->fd, &f, sizeof(f)) || !recv_frame(r->fd, &rep)) return B200KV_EIO;
  if (rep.status != B200KV_OK) return rep.status;
  uint32_t slot = 0;
  int rc = rep.length == sb && rep.n_tokens > 0 ? b200kv_pool_reserve(local, key, rep.n_tokens, rep.fmt, owner, &slot)
                                                : B200KV_EINVAL;
  if (rc != B200KV_OK) {
    if (!drain(r->fd, rep.length)) return B200KV_EIO;
    return rc == B200KV_EEXIST ? B200KV_OK : rc;
  }
  if (!recv_all(r->fd, b200kv_pool_slot_ptr(local, slot), sb)) {
    b200kv_pool_abort(local, key);
    return B200KV_EIO;
  }
  r->bytes_down += sb;
  return b200kv_pool_commit(local, key);
}

extern "C" int b200kv_remote_stats(b200kv_remote* r, b200kv_pool_stats* out) {
  if (!r || !out) return B200KV_EINVAL;
  std::lock_guard<std::mutex> lk(r->mu);
  Frame f = make_frame(kStats), rep;
  if (!send_all(r->fd, &f, sizeof(f)) || !recv_frame(r->fd, &rep) || rep.length != sizeof(*out) ||
      !recv_all(r->fd, out, sizeof(*out)))
    return B200KV_EIO;
  return rep.status;
}

extern "C" int b200kv_remote_traffic(b200kv_remote* r, uint64_t* bytes_up, uint64_t* bytes_down) {
  if (!r) return B200KV_EINVAL;
  std::lock_guard<std::mutex> lk(r->mu);
  if (bytes_up) *bytes_up = r->bytes_up;
  if (bytes_down) *bytes_down = r->bytes_down;
  return B200KV_OK;
}
