"""Debug helper (GPU box): render goldens, print differing pixels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package, ASSETS
from conftest import load_golden
pkg = load_package()
tags = sys.argv[1:] or ["c1_500_spp128", "c1_500_spp16", "c1_200_spp64", "ao_500_spp128", "tri_500_spp128", "teapots_160x90_spp16", "teapots_96_spp1", "teapots_point_96x64_spp4", "mix_small_128x72_spp4", "wide_37x23_spp3", "c1_200_spp8_d0", "c1_200_spp8_d2"]
for tag in tags:
    g = load_golden(tag)
    for trav in (pkg.TRAVERSAL_AUTO, pkg.TRAVERSAL_BVH):
        rt = pkg.Raytracer(g["W"], g["H"]); rt.SetAssetsPath(ASSETS)
        rt.SetOptions(depth=g["depth"], ao_spp=g["spp"], traversal=trav)
        assert rt.LoadSceneJSON(g["scene"]) == 0
        st = rt.Render("")
        if st != 0:
            print(tag, trav, "RENDER FAILED", pkg.lib().rt580_last_error()); continue
        fb = rt.frame_buffer(); s = rt.stats()
        d = (fb != g["fb"]).any(axis=-1)
        print("%-28s trav=%d rays %d vs %d  diff pixels %d  ms_total %.3f (struct %.3f order %.3f ao %.3f resolve %.3f) launches %d" % (
            tag, trav, s.rays, g["rays"], int(d.sum()), s.ms_total, s.ms_structure, s.ms_order, s.ms_ao, s.ms_resolve, s.kernel_launches))
        ys, xs = np.nonzero(d)
        for y, x in list(zip(ys, xs))[:12]:
            print("   (%d,%d) got %s want %s" % (x, y, fb[y, x], g["fb"][y, x]))
