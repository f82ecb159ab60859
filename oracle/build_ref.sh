#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle tier T0) - builds oracle/_ref/libref580.so from
# the reference's own sources WHERE THEY LIE under /root/reference.  No reference
# source is copied into the repo: both files are streamed through sed straight into
# g++ (stdin), and only the resulting .so lands in the git-ignored oracle/_ref/.
#
# The reference is MSVC-only as written; the stream applies these mechanical,
# arithmetic-neutral edits (all verified to leave the default
# 500x500 image byte-identical, md5 a00a8b5cb0a0e7dd3bd85f675a43e94a):
#   Raytracer.h:18-34   delete the private forward declarations (g++: "redeclared
#                       with different access")
#   Raytracer.h:15      ASSETS_PATH loses `const` so fixtures can live outside CWD
#   Raytracer.cpp:1,3,4 drop the includes (header is streamed in front; json.hpp is
#                       pre-included; CImg is never used and needs X11)
#   Raytracer.cpp:253   std::powf -> ::powf via a using-declaration shim
#   Raytracer.cpp:317   128 -> g_ref_spp
#   Raytracer.cpp:925   Raycast(ray) -> Raycast(ray, g_ref_depth)
#   Raytracer.cpp:926-927 per-pixel progress print removed
#   Raytracer.cpp:473   IntersectScene counts its calls in g_ref_rays ("1 ray")
#   Raytracer.cpp:944   main renamed
# Compile flags: -O3 -march=x86-64-v3 (portable to the GPU box host) -ffp-contract=off (no FMA contraction: the
# reference's MSVC /fp:precise build has none either; survey measured that allowing
# contraction changes 32 pixels).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF580_SRC:-/root/reference/580 Raytracer}"
out="$here/_ref"
if [ ! -f "$REF/Raytracer.cpp" ]; then
    echo "build_ref.sh: reference sources not present at $REF (expected on the GPU box); keeping prebuilt $out" >&2
    exit 0
fi
mkdir -p "$out"
{
    cat <<'EOF'
#include <cmath>
#include <vector>
#include <unordered_map>
#include <string>
#include <random>
#include <iostream>
#include <fstream>
#include <chrono>
#include "ExternalPlugins/json.hpp"
namespace std { using ::powf; }
static thread_local unsigned long long g_ref_rays = 0;
static int g_ref_spp = 128;
static int g_ref_depth = 4;
#define private public
EOF
    sed -e '1s/^\xEF\xBB\xBF//' -e '18,34d' \
        -e 's/^const std::string ASSETS_PATH/std::string ASSETS_PATH/' \
        "$REF/Raytracer.h"
    sed -e '1s/^\xEF\xBB\xBF//' -e 's/^#include "Raytracer.h"//' \
        -e 's/^#include "ExternalPlugins\/json.hpp"//' \
        -e 's/^#include "ExternalPlugins\/CImg\/CImg.h"//' \
        -e 's/int numSamples = 128;/int numSamples = g_ref_spp;/' \
        -e 's/= Raycast(ray);/= Raycast(ray, g_ref_depth);/' \
        -e '/Rendered: /d' \
        -e '/std::cout.flush();/d' \
        -e 's/^bool Raytracer::IntersectScene(const Ray& ray, RaycastHitInfo& hitInfo) {/&\n\tg_ref_rays++;/' \
        -e 's/^int main() {/int ref580_unused_main() {/' \
        "$REF/Raytracer.cpp"
    printf '\n#undef private\n'
    cat "$here/ref_driver.inc"
} | g++ -std=c++17 -O3 -march=x86-64-v3 -ffp-contract=off -fPIC -shared -pthread -w \
        -I"$REF" -x c++ - -o "$out/libref580.so"
echo "built $out/libref580.so"
