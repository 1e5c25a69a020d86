// Procedural test mesh shared by the host-side debug tools (tools/host_probe.cu, tools/coop_check.cu).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

// bumpy icosphere identical in spirit to zombie_b200/procedural.py (not bit-identical: own RNG)
static void icosphere(int subdiv, std::vector<float>& pos, std::vector<int>& tri)
{
    const double t = (1.0 + std::sqrt(5.0))/2.0;
    double v0[12][3] = {{-1,t,0},{1,t,0},{-1,-t,0},{1,-t,0},{0,-1,t},{0,1,t},{0,-1,-t},{0,1,-t},{t,0,-1},{t,0,1},{-t,0,-1},{-t,0,1}};
    int f0[20][3] = {{0,11,5},{0,5,1},{0,1,7},{0,7,10},{0,10,11},{1,5,9},{5,11,4},{11,10,2},{10,7,6},{7,1,8},{3,9,4},{3,4,2},{3,2,6},{3,6,8},{3,8,9},{4,9,5},{2,4,11},{6,2,10},{8,6,7},{9,8,1}};
    std::vector<double> p;
    for (auto& v : v0) { double l = std::sqrt(v[0]*v[0]+v[1]*v[1]+v[2]*v[2]); p.push_back(v[0]/l); p.push_back(v[1]/l); p.push_back(v[2]/l); }
    std::vector<int> f;
    for (auto& ff : f0) { f.push_back(ff[0]); f.push_back(ff[1]); f.push_back(ff[2]); }
    for (int s = 0; s < subdiv; s++) {
        std::map<std::pair<int,int>, int> mid;
        std::vector<int> nf;
        auto midpoint = [&](int a, int b) {
            auto key = std::make_pair(std::min(a,b), std::max(a,b));
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            double m[3] = {p[3*a]+p[3*b], p[3*a+1]+p[3*b+1], p[3*a+2]+p[3*b+2]};
            double l = std::sqrt(m[0]*m[0]+m[1]*m[1]+m[2]*m[2]);
            int idx = (int)p.size()/3;
            p.push_back(m[0]/l); p.push_back(m[1]/l); p.push_back(m[2]/l);
            mid[key] = idx;
            return idx;
        };
        for (size_t i = 0; i < f.size(); i += 3) {
            int a = f[i], b = f[i+1], c = f[i+2];
            int ab = midpoint(a,b), bc = midpoint(b,c), ca = midpoint(c,a);
            int t4[12] = {a,ab,ca, b,bc,ab, c,ca,bc, ab,bc,ca};
            nf.insert(nf.end(), t4, t4+12);
        }
        f.swap(nf);
    }
    // bumps
    double w[3][3] = {{1.3,-2.1,0.7},{-3.2,0.4,2.5},{0.9,3.1,-1.7}}, ph[3] = {0.3, 1.7, 4.1};
    pos.resize(p.size());
    for (size_t i = 0; i < p.size(); i += 3) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += std::sin(w[k][0]*p[i]+w[k][1]*p[i+1]+w[k][2]*p[i+2]+ph[k]);
        double r = 1.0 + 0.15*s/3.0;
        pos[i] = (float)(p[i]*r); pos[i+1] = (float)(p[i+1]*r); pos[i+2] = (float)(p[i+2]*r);
    }
    tri = f;
}

