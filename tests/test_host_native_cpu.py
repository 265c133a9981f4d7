"""Host-side native pieces of the ingest compiled on their own with g++ and checked on the CPU: the DEFLATE decoder
(csrc/inflate_fast.hpp) against zlib on random streams of every block type, with guard bytes around the output and
truncated / corrupted / mis-sized inputs; the CRC-32 by carry-less multiplication (csrc/crc32_clmul.hpp) against zlib's;
the open-addressing name table of the admission pass (csrc/overlap.hpp) against std::unordered_map."""
import os
import subprocess

import pytest

from helpers import ROOT

CSRC = os.path.join(ROOT, "covid-spings-variant-caller_b200", "csrc")

HARNESS = r"""
#include "inflate_fast.hpp"
#include "crc32_clmul.hpp"
#include "overlap.hpp"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>
using namespace std;

static vector<uint8_t> deflate_raw(const vector<uint8_t>& in, int level, int strategy) {
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
    zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size(); zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { printf("deflate failed\n"); exit(2); }
    out.resize(zs.total_out); deflateEnd(&zs);
    return out;
}
static bool guards(const vector<uint8_t>& o, size_t n) {
    for (int g = 0; g < 8; ++g) if (o[g] != 0xEE || o[8 + n + g] != 0xEE) return false;
    return true;
}

static long test_inflate() {
    lvc_inflate::Tables* T = new lvc_inflate::Tables();
    srand(12345);
    long fails = 0;
    for (int iter = 0; iter < 700; ++iter) {
        const size_t n = iter < 40 ? (size_t)iter + 1 : (size_t)(rand() % 65536) + 1;
        vector<uint8_t> data(n);
        const int kind = rand() % 7;
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
                case 0: data[i] = (uint8_t)rand(); break;
                case 1: data[i] = (uint8_t)"ACGT"[rand() & 3]; break;
                case 2: data[i] = (i > 300 && (rand() % 50)) ? data[i - 1 - rand() % 300] : (uint8_t)rand(); break;
                case 3: data[i] = (uint8_t)(i / 7); break;
                case 4: data[i] = (rand() % 20) ? (uint8_t)'A' : (uint8_t)rand(); break;
                case 5: data[i] = (i >= 3 && (rand() % 8)) ? data[i - 3] : (uint8_t)(rand() % 5 + 60); break;
                default: { int r = rand(), v = 0; while ((r & 1) && v < 40) { r >>= 1; ++v; } data[i] = (uint8_t)(v * 5 + (rand() % 3 == 0)); }   // skewed: long codes
            }
        }
        const int level = rand() % 10;
        const int strat = (rand() % 5 == 0) ? Z_FIXED : ((rand() % 7 == 0) ? Z_HUFFMAN_ONLY : ((rand() % 9 == 0) ? Z_RLE : Z_DEFAULT_STRATEGY));
        const vector<uint8_t> c = deflate_raw(data, level, strat);
        for (int slack : {0, 8, 16, 40}) {
            vector<uint8_t> in(c.size() + (size_t)slack, 0xA5);
            memcpy(in.data(), c.data(), c.size());
            vector<uint8_t> out(n + 16, 0xEE);
            const bool ok = lvc_inflate::inflate_block(*T, in.data(), c.size(), (size_t)slack, out.data() + 8, n);
            if (!ok || memcmp(out.data() + 8, data.data(), n) != 0 || !guards(out, n)) { ++fails; if (fails < 6) printf("inflate iter %d n %zu kind %d level %d strat %d slack %d ok %d\n", iter, n, kind, level, strat, slack, (int)ok); }
            if (n > 1) {                                     // a smaller announced size: rejected, nothing written past it
                vector<uint8_t> o2(n + 16, 0xEE);
                if (lvc_inflate::inflate_block(*T, in.data(), c.size(), (size_t)slack, o2.data() + 8, n - 1) || !guards(o2, n - 1)) ++fails;
            }
            if (c.size() > 4) {                              // a truncated stream: rejected (or, cut inside its padding, still right)
                vector<uint8_t> o3(n + 16, 0xEE);
                const size_t cut = c.size() - 1 - (size_t)(rand() % 3);
                const bool ok3 = lvc_inflate::inflate_block(*T, in.data(), cut, (size_t)slack + (c.size() - cut), o3.data() + 8, n);
                if ((ok3 && memcmp(o3.data() + 8, data.data(), n) != 0) || !guards(o3, n)) ++fails;
            }
        }
        for (int t = 0; t < 3 && c.size() > 2; ++t) {        // a flipped bit: any verdict, but never a write outside the buffer
            vector<uint8_t> in(c.size() + 16, 0x5A);
            memcpy(in.data(), c.data(), c.size());
            in[(size_t)rand() % c.size()] ^= (uint8_t)(1 << (rand() % 8));
            vector<uint8_t> o4(n + 16, 0xEE);
            lvc_inflate::inflate_block(*T, in.data(), c.size(), 16, o4.data() + 8, n);
            if (!guards(o4, n)) ++fails;
        }
    }
    delete T;
    return fails;
}

static long test_crc() {
    srand(7);
    long fails = 0;
    vector<uint8_t> b(70000);
    for (auto& x : b) x = (uint8_t)rand();
    for (int it = 0; it < 4000; ++it) {
        const size_t n = it < 300 ? (size_t)it : (size_t)(rand() % 65537), off = (size_t)(rand() % 64);
        if (lvc_crc::crc32_block(b.data() + off, n) != (uint32_t)crc32(crc32(0, 0, 0), b.data() + off, (uInt)n)) ++fails;
    }
    return fails;
}

static long test_name_table() {
    srand(99);
    long fails = 0;
    for (int round = 0; round < 60; ++round) {
        lvc_overlap::NameTable t;
        unordered_map<uint64_t, uint32_t> ref;
        const int space = 1 + rand() % 3000;
        const uint64_t mult = (round % 3 == 0) ? 1024 : 1;   // many keys with the same home slot
        for (int k = 0; k < 20000; ++k) {
            const uint64_t h = (uint64_t)(rand() % space) * mult + (round % 5 == 0 ? 0 : 7);
            const long at = t.find(h, [&](uint32_t j) { return ref.count(h) && ref[h] == j; });
            if ((at >= 0) != (ref.count(h) != 0)) ++fails;
            if (at >= 0) {
                if (t.t[(size_t)at].idx1 - 1 != ref[h]) ++fails;
                if (rand() % 3) { t.erase(at); ref.erase(h); }
            } else if (rand() % 2) { t.insert(h, (uint32_t)k); ref[h] = (uint32_t)k; }
            if (t.count != ref.size()) ++fails;
        }
        for (auto& kv : ref) if (t.find(kv.first, [&](uint32_t j) { return j == kv.second; }) < 0) ++fails;
    }
    return fails;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "";
    long f = -1;
    if (!strcmp(what, "inflate")) f = test_inflate();
    else if (!strcmp(what, "crc")) f = test_crc();
    else if (!strcmp(what, "names")) f = test_name_table();
    printf("%s fails %ld clmul %d\n", what, f, (int)lvc_crc::clmul_usable());
    return f == 0 ? 0 : 1;
}
"""


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("native")
    src, exe = os.path.join(d, "h.cpp"), os.path.join(d, "h")
    with open(src, "w") as fh:
        fh.write(HARNESS)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-pthread",
                    "-I", CSRC, "-o", exe, src, "-lz"], check=True)
    return exe


@pytest.mark.parametrize("what", ["inflate", "crc", "names"])
def test_native_piece(harness, what):
    r = subprocess.run([harness, what], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"{what} fails 0" in r.stdout
