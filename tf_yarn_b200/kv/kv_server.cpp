// tfy_kv: the key-value rendezvous server of the local launcher.
//
// Stand-in for the skein ApplicationMaster's gRPC key-value store that the
// reference uses as its whole control plane (reference: tf_yarn/event.py:13-18,
// 70-79 -- kv.wait / kv[key]=value; tf_yarn/client.py:566-568,647 -- keys(),
// events("PUT")).  Verbs: PUT, GET, WAIT (blocking until the key exists), KEYS
// (by prefix), DEL, SUBSCRIBE (stream of every PUT, replaying the existing
// keys first so a late subscriber misses nothing), PING.
//
// Wire format (little endian), both directions:
//     u32 frame_len | u8 op_or_status | u32 klen | key | u32 vlen | value
// One thread per connection; a WAIT parks its connection thread on a condition
// variable, so a waiting task costs no polling.  The store carries ~100 small
// keys per job (addresses, lifecycle events, the pickled experiment), it is not
// on any data path.
#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

enum Op : uint8_t { OP_PUT = 1, OP_GET = 2, OP_WAIT = 3, OP_KEYS = 4, OP_SUBSCRIBE = 5, OP_DEL = 6, OP_PING = 7 };
enum Status : uint8_t { ST_OK = 0, ST_NOTFOUND = 1, ST_EVENT = 2, ST_ERROR = 3 };

bool read_all(int fd, void* buf, size_t n) {
    char* p = (char*)buf;
    while (n) {
        ssize_t r = recv(fd, p, n, 0);
        if (r == 0) return false;
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        n -= (size_t)r;
    }
    return true;
}

bool write_all(int fd, const void* buf, size_t n) {
    const char* p = (const char*)buf;
    while (n) {
        ssize_t r = send(fd, p, n, MSG_NOSIGNAL);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        n -= (size_t)r;
    }
    return true;
}

bool send_frame(int fd, uint8_t status, const std::string& key, const std::string& val) {
    uint32_t klen = (uint32_t)key.size(), vlen = (uint32_t)val.size();
    uint32_t len = 1 + 4 + klen + 4 + vlen;
    std::string out;
    out.reserve(4 + len);
    out.append((const char*)&len, 4);
    out.push_back((char)status);
    out.append((const char*)&klen, 4);
    out.append(key);
    out.append((const char*)&vlen, 4);
    out.append(val);
    return write_all(fd, out.data(), out.size());
}

struct Server {
    int listen_fd = -1;
    int port = 0;
    std::atomic<bool> stop{false};
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::string, std::string> store;
    uint64_t version = 0;                                       // bumps on every PUT
    std::deque<std::pair<std::string, std::string>> log;        // ordered PUT log for subscribers
    std::thread accept_thread;
    struct Conn {
        int fd = -1;                       // -1 once serve() has closed it
        std::atomic<bool> done{false};     // serve() returned: the thread can be joined without blocking
        std::thread th;
    };
    std::mutex conn_mu;
    std::list<std::unique_ptr<Conn>> conns;

    // A finished connection releases its descriptor immediately (KVClient.wait() opens one connection per call, a
    // launcher sees hundreds of them per job); the std::thread object is joined and dropped by the next accept.
    void finish(Conn* c) {
        std::lock_guard<std::mutex> g(conn_mu);
        if (c->fd >= 0) {
            ::shutdown(c->fd, SHUT_RDWR);
            close(c->fd);
            c->fd = -1;
        }
        c->done.store(true);
    }

    void reap_locked() {
        for (auto it = conns.begin(); it != conns.end();) {
            if ((*it)->done.load()) {
                if ((*it)->th.joinable()) (*it)->th.join();
                it = conns.erase(it);
            } else {
                ++it;
            }
        }
    }

    void serve(int fd) {
        int one = 1;
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        for (;;) {
            uint32_t len = 0;
            if (!read_all(fd, &len, 4) || len < 9 || len > (1u << 30)) break;
            std::string frame(len, '\0');
            if (!read_all(fd, &frame[0], len)) break;
            uint8_t op = (uint8_t)frame[0];
            uint32_t klen = 0;
            memcpy(&klen, &frame[1], 4);
            if (5ull + klen + 4 > len) break;
            std::string key = frame.substr(5, klen);
            uint32_t vlen = 0;
            memcpy(&vlen, &frame[5 + klen], 4);
            if (9ull + klen + vlen > len) break;
            std::string val = frame.substr(9 + klen, vlen);
            bool ok = true;
            switch (op) {
                case OP_PUT: {
                    {
                        std::lock_guard<std::mutex> g(mu);
                        store[key] = val;
                        log.emplace_back(key, val);
                        ++version;
                    }
                    cv.notify_all();
                    ok = send_frame(fd, ST_OK, "", "");
                    break;
                }
                case OP_GET: {
                    std::unique_lock<std::mutex> g(mu);
                    auto it = store.find(key);
                    if (it == store.end()) {
                        g.unlock();
                        ok = send_frame(fd, ST_NOTFOUND, key, "");
                    } else {
                        std::string v = it->second;
                        g.unlock();
                        ok = send_frame(fd, ST_OK, key, v);
                    }
                    break;
                }
                case OP_WAIT: {
                    std::unique_lock<std::mutex> g(mu);
                    cv.wait(g, [&] { return stop.load() || store.count(key) > 0; });
                    if (stop.load() && !store.count(key)) {
                        g.unlock();
                        ok = send_frame(fd, ST_ERROR, key, "server stopping");
                    } else {
                        std::string v = store[key];
                        g.unlock();
                        ok = send_frame(fd, ST_OK, key, v);
                    }
                    break;
                }
                case OP_KEYS: {
                    std::string joined;
                    {
                        std::lock_guard<std::mutex> g(mu);
                        for (auto it = store.lower_bound(key); it != store.end(); ++it) {
                            if (it->first.compare(0, key.size(), key) != 0) break;
                            joined.append(it->first);
                            joined.push_back('\n');
                        }
                    }
                    ok = send_frame(fd, ST_OK, key, joined);
                    break;
                }
                case OP_DEL: {
                    {
                        std::lock_guard<std::mutex> g(mu);
                        store.erase(key);
                    }
                    ok = send_frame(fd, ST_OK, "", "");
                    break;
                }
                case OP_PING: {
                    ok = send_frame(fd, ST_OK, "", "pong");
                    break;
                }
                case OP_SUBSCRIBE: {
                    // stream the PUT log from the beginning; never returns to request mode
                    size_t cursor = 0;
                    for (;;) {
                        std::vector<std::pair<std::string, std::string>> batch;
                        {
                            std::unique_lock<std::mutex> g(mu);
                            cv.wait(g, [&] { return stop.load() || log.size() > cursor; });
                            if (stop.load() && log.size() <= cursor) { ok = false; break; }
                            while (cursor < log.size()) batch.push_back(log[cursor++]);
                        }
                        for (auto& kv : batch)
                            if (!send_frame(fd, ST_EVENT, kv.first, kv.second)) { ok = false; break; }
                        if (!ok) break;
                    }
                    break;
                }
                default:
                    ok = send_frame(fd, ST_ERROR, "", "bad op");
            }
            if (!ok) break;
        }
    }

    void accept_loop() {
        for (;;) {
            int fd = accept(listen_fd, nullptr, nullptr);
            if (fd < 0) {
                if (stop.load()) return;
                if (errno == EINTR) continue;
                if (errno == EMFILE || errno == ENFILE || errno == ENOBUFS || errno == ENOMEM ||
                    errno == ECONNABORTED || errno == EPROTO || errno == EAGAIN) {
                    // transient: out of descriptors / aborted handshake.  Keep serving -- a store that silently
                    // stops accepting hangs every task in kv.wait().
                    fprintf(stderr, "[tfy_kv] accept: %s; retrying\n", strerror(errno));
                    {
                        std::lock_guard<std::mutex> g(conn_mu);
                        reap_locked();
                    }
                    usleep(20000);
                    continue;
                }
                fprintf(stderr, "[tfy_kv] accept failed permanently: %s\n", strerror(errno));
                return;
            }
            std::lock_guard<std::mutex> g(conn_mu);
            if (stop.load()) { close(fd); return; }
            reap_locked();
            conns.emplace_back(new Conn());
            Conn* c = conns.back().get();
            c->fd = fd;
            c->th = std::thread([this, c, fd] {
                serve(fd);
                finish(c);
            });
        }
    }
};

}  // namespace

extern "C" {

void* tfy_kv_start(const char* host, int port) {
    auto* s = new Server();
    s->listen_fd = socket(AF_INET, SOCK_STREAM, 0);
    if (s->listen_fd < 0) { delete s; return nullptr; }
    int one = 1;
    setsockopt(s->listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    if (!host || !*host || inet_pton(AF_INET, host, &addr.sin_addr) != 1) addr.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    if (bind(s->listen_fd, (sockaddr*)&addr, sizeof(addr)) != 0 || listen(s->listen_fd, 256) != 0) {
        close(s->listen_fd);
        delete s;
        return nullptr;
    }
    socklen_t alen = sizeof(addr);
    getsockname(s->listen_fd, (sockaddr*)&addr, &alen);
    s->port = ntohs(addr.sin_port);
    s->accept_thread = std::thread([s] { s->accept_loop(); });
    return s;
}

int tfy_kv_port(void* h) { return h ? ((Server*)h)->port : -1; }

void tfy_kv_stop(void* h) {
    auto* s = (Server*)h;
    if (!s) return;
    s->stop.store(true);
    {
        std::lock_guard<std::mutex> g(s->mu);
    }
    s->cv.notify_all();
    ::shutdown(s->listen_fd, SHUT_RDWR);
    close(s->listen_fd);
    if (s->accept_thread.joinable()) s->accept_thread.join();
    {
        std::lock_guard<std::mutex> g(s->conn_mu);
        for (auto& c : s->conns)
            if (c->fd >= 0) ::shutdown(c->fd, SHUT_RDWR);      // wakes a thread parked in recv()
    }
    for (auto& c : s->conns)                                   // finish() closes the descriptors
        if (c->th.joinable()) c->th.join();
    delete s;
}

}  // extern "C"
