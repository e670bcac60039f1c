// TEST INFRASTRUCTURE - NOT PRODUCT CODE, NOT REFERENCE CODE.
//
// Drop-in replacement TU for /root/reference/ECMAgentSimulator/KDTree.cpp implementing the SAME
// class (KDTree.h:57-80) with the *exact* k-nearest specification the GPU neighbour kernel is
// held to (SURVEY.md §8c "exact-knn"):
//   candidates = agents active when Construct() ran (KDTree.cpp:34-41);
//   sqDist     = fl(fl(dx*dx) + fl(dy*dy)), dx = pos[j].x - pos[i].x   (KDTree.cpp:106-108);
//   keep        sqDist > EPSILON (1e-4f)                                (KDTree.cpp:112, 146);
//   result     = the k smallest by (sqDist, slot index) ascending, in that order;
//   count      = min(k, #candidates kept).   Unused output slots are set to -1.
// The reference's own query is an approximate kNN; linking this TU instead
// of KDTree.cpp changes nothing else in the reference build.
//
// Implementation: candidates sorted by x; the scan walks outwards from the query's rank and stops
// on a side once dx*dx exceeds the current k-th distance.  m_Tree holds the x-sorted slot list.
#include "KDTree.h"

#include "Simulator.h"
#include "ECMDataTypes.h"
#include "UtilityFunctions.h"

namespace ECM {
namespace Simulation {

int KDTree::LeftTree(int root) const { return root * 2 + 1; }
int KDTree::RightTree(int root) const { return root * 2 + 2; }

void KDTree::Construct(Simulator* simulation) {
    PositionComponent* positions = simulation->GetPositionData();
    bool* activeFlags = simulation->GetActiveFlags();
    m_Tree.clear();
    for (int i = 0; i <= simulation->GetLastIndex(); i++)
        if (activeFlags[i]) m_Tree.push_back(i);
    std::sort(m_Tree.begin(), m_Tree.end(), [positions](int a, int b) {
        return positions[a].x < positions[b].x || (positions[a].x == positions[b].x && a < b);
    });
    m_MaxDepth = 0;
}

void KDTree::ConstructRecursive(PositionComponent*, int, KDTreeCompareY&, KDTreeCompareX&, int, int, int, int*) {}
void KDTree::KNearestAgents_R(const Vec2&, int, int, int&, int, std::vector<Entity>&, std::vector<float>&, PositionComponent*) {}
void KDTree::AgentsInRange_R(const Vec2&, float, int, int, int&, int, PositionComponent*, std::vector<int>&) {}
void KDTree::AgentsInRange(Simulator*, int, float, std::vector<Entity>&, int&) {}
void KDTree::AgentsInRangeTest(PositionComponent*, int, float, int, std::vector<int>&) {}
void KDTree::TestConstruct(PositionComponent*, int) {}
void KDTree::Clear() { m_Tree.clear(); }

void KDTree::KNearestAgents(Simulator* simulation, int agent, int k, std::vector<Entity>& outAgents, int& outNumNeighbors) {
    PositionComponent* positions = simulation->GetPositionData();
    const float tx = positions[agent].x, ty = positions[agent].y;
    const int n = (int)m_Tree.size();
    std::vector<float> best(k, Utility::MAX_FLOAT);
    std::vector<int> ids(k, -1);
    int found = 0;
    auto consider = [&](int j) {
        const float dx = positions[j].x - tx, dy = positions[j].y - ty;
        const float mx = dx * dx, my = dy * dy;
        const float d = mx + my;
        if (!(d > Utility::EPSILON)) return;
        int p = found < k ? found : k;  // insertion position search from the back
        while (p > 0 && (d < best[p - 1] || (d == best[p - 1] && j < ids[p - 1]))) p--;
        if (p >= k) return;
        for (int q = (found < k ? found : k - 1); q > p; q--) {
            best[q] = best[q - 1];
            ids[q] = ids[q - 1];
        }
        best[p] = d;
        ids[p] = j;
        if (found < k) found++;
    };
    // rank of the query position in the x-sorted list
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (positions[m_Tree[mid]].x < tx) lo = mid + 1; else hi = mid;
    }
    int r = lo, l = lo - 1;
    bool goR = true, goL = true;
    while (goR || goL) {
        if (goR) {
            if (r >= n) goR = false;
            else {
                const float dx = positions[m_Tree[r]].x - tx;
                if (found == k && dx * dx > best[k - 1]) goR = false;
                else consider(m_Tree[r++]);
            }
        }
        if (goL) {
            if (l < 0) goL = false;
            else {
                const float dx = positions[m_Tree[l]].x - tx;
                if (found == k && dx * dx > best[k - 1]) goL = false;
                else consider(m_Tree[l--]);
            }
        }
    }
    outNumNeighbors = found;
    for (int i = 0; i < k; i++) outAgents[i] = ids[i];
}

}  // namespace Simulation
}  // namespace ECM
