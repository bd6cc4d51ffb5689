#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cstdint>
extern "C" {
void *b200host_chain_new2(int, unsigned, unsigned, unsigned, unsigned);
int b200host_chain_run(void *, const double *, long, long, int, uint8_t *, int, int32_t *, int, double *);
void b200host_chain_delete(void *);
double *b200host_alloc_samples(long);
void b200host_free_samples(double *);
void *b200host_rx_block_new2(int, unsigned, unsigned, unsigned, unsigned);
int b200host_rx_block_run(void *, const double *, const uint8_t *, long, long, uint8_t *, int, int32_t *, int, double *);
void b200host_rx_block_delete(void *);
int b200host_receiver_run(const double *, long, long, long, int, unsigned, unsigned, uint8_t *, int, int32_t *, int, long *);
}
template <class T> std::vector<T> slurp(const char *p) { FILE *f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); rewind(f); std::vector<T> v(n / sizeof(T)); if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) abort(); fclose(f); return v; }
int main() {
    auto x = slurp<double>("/tmp/asan/stream.bin");
    auto s = slurp<double>("/tmp/asan/synced.bin");
    auto t = slurp<uint8_t>("/tmp/asan/tags.bin");
    std::vector<uint8_t> pl(512 * 4095); std::vector<int32_t> ln(512); double sec[2];
    long chunks[] = {333, 4096, 20000, 70000, (long)(x.size() / 2)};
    for (long c : chunks) for (unsigned depth : {1u, 3u, 6u}) {
        void *ch = b200host_chain_new2(0, 64, 4095, depth, depth - 1);
        if (!ch) { puts("chain_new failed"); return 1; }
        int n1 = b200host_chain_run(ch, x.data(), x.size() / 2, c, c % 2, pl.data(), 4095, ln.data(), 512, sec);
        int n2 = b200host_chain_run(ch, x.data(), x.size() / 2, c, 0, pl.data(), 4095, ln.data(), 512, sec);
        b200host_chain_delete(ch);
        printf("chain chunk %ld depth %u: %d %d payloads\n", c, depth, n1, n2);
    }
    for (long c : {4096L, 30000L}) for (unsigned depth : {1u, 4u}) {
        void *b = b200host_rx_block_new2(0, 16, 4095, depth, depth - 1);
        int n1 = b200host_rx_block_run(b, s.data(), t.data(), (long)t.size(), c, pl.data(), 4095, ln.data(), 512, sec);
        b200host_rx_block_delete(b);
        printf("block round %ld depth %u: %d payloads\n", c, depth, n1);
    }
    {   // pinned caller buffer: long calls go to the "GPU" straight from it
        double *pin = b200host_alloc_samples((long)(x.size() / 2));
        for (size_t i = 0; i < x.size(); i++) pin[i] = x[i];
        for (long c : {65536L, 70001L, 200000L}) for (unsigned depth : {1u, 6u}) {
            void *ch = b200host_chain_new2(0, 64, 4095, depth, depth - 1);
            int n1 = b200host_chain_run(ch, pin, x.size() / 2, c, 0, pl.data(), 4095, ln.data(), 512, sec);
            b200host_chain_delete(ch);
            printf("pinned chain chunk %ld depth %u: %d payloads\n", c, depth, n1);
        }
        b200host_free_samples(pin);
    }
    long paused = -1;
    int n3 = b200host_receiver_run(x.data(), x.size() / 2, 4096, 3, 5, 64, 4095, pl.data(), 4095, ln.data(), 512, &paused);
    printf("receiver: %d payloads, rounds while paused %ld\n", n3, paused);
    return 0;
}
