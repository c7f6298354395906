// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, no dependencies) of the read-mapping hot path of
// gFedonin/VirGenA, written from the Java sources. Every function cites the
// reference file:line it follows. This file is the parity checker for the CUDA
// path and the "port" CPU baseline of bench.py; it is never linked into, nor
// called from, the product library (virgena_b200/csrc). Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.
//
// PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, fixtures or
// golden vectors for this path and cannot be compiled or
// run in this environment (no JVM, jars not vendored). What pins this oracle
// instead are hand-derived known-answer tests in tests/test_oracle_kat.py (one
// per quirk Q1..Q12 of SURVEY.md §8a) worked out on paper from the Java source.
//
// Java semantics that are emulated on purpose:
//   * TreeSet<LongHit> ordered by (genomePos, readPos); add() of an equal key is
//     a no-op that keeps the element already present (Hit.java:13-19).
//   * ArrayList.sort / Collections.sort are stable.
//   * int division truncates toward zero; `x*3/2` is `(x*3)/2`.
//   * concordanceArray divides in IEEE-754 binary32 (KMerCounterBase.java:272).
//   * SmithWatermanGotoh keeps scores in `float` holding exact small integers;
//     they are kept here as int with a large negative sentinel for -inf that is
//     consumed only in row 1 / column 1 exactly as in the Java.

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace vo {

// ----------------------------------------------------------------------------
// Constants.java
// ----------------------------------------------------------------------------
static const uint8_t GAP = '-';

struct Tables {
  uint8_t complement[127];
  uint8_t lower[127];
  int nToI['T' + 1];
  uint8_t iToN[5];
  Tables() {
    memset(complement, 0, sizeof complement);  // Constants.java:11-20
    complement['A'] = 'T';
    complement['T'] = 'A';
    complement['G'] = 'C';
    complement['C'] = 'G';
    complement[GAP] = GAP;
    complement['N'] = 'N';
    memset(lower, 0, sizeof lower);  // Constants.java:22-30
    lower['A'] = 'a';
    lower['T'] = 't';
    lower['G'] = 'g';
    lower['C'] = 'c';
    lower['N'] = 'n';
    for (int i = 0; i <= 'T'; i++) nToI[i] = 0;  // Constants.java:51-60 (Q11)
    nToI['A'] = 0;
    nToI['T'] = 1;
    nToI['G'] = 2;
    nToI['C'] = 3;
    nToI[GAP] = 4;
    iToN[0] = 'A';
    iToN[1] = 'T';
    iToN[2] = 'G';
    iToN[3] = 'C';
    iToN[4] = GAP;
  }
};
static const Tables T;

// Constants.getComplement (Constants.java:42-49). Bytes >= 127 would throw in
// Java; callers validate the alphabet first.
static std::string getComplement(const std::string& seq) {
  size_t len = seq.size();
  std::string res(len, '\0');
  for (size_t i = 0; i < len; i++) res[i] = (char)T.complement[(uint8_t)seq[len - i - 1]];
  return res;
}

// ----------------------------------------------------------------------------
// Types: Hit.java, LongHit.java, Location.java, MappedRead.java, Alignment.java
// ----------------------------------------------------------------------------
struct LongHit {
  int genomePos, readPos, len;
};

struct Location {
  int start = 0, end = 0, startIndex = 0, endIndex = 0, count = 0;
};

struct Alignment {
  int score = 0;
  std::string sequence1;
  int start1 = 0, end1 = 0, start2 = 0, end2 = 0, identity = 0, length = 0;
};

struct MappedRead {
  int start = 0, end = 0, count = 0;
  int reverse = 0;
  int fragmentID = 0;
  bool hasAln = false;
  Alignment aln;
  // MappedRead.chooseFragment (MappedRead.java:39-50); Q7: fragmentID stays 0
  // when the window fits no fragment.
  void chooseFragment(const std::vector<int>& fragmentEnds) {
    int s = 0, e = 0;
    for (size_t i = 0; i < fragmentEnds.size(); i++) {
      e = fragmentEnds[i];
      if (start >= s && end <= e) {
        fragmentID = (int)i;
        break;
      }
      s = e;
    }
  }
};

typedef std::unordered_map<std::string, std::vector<int>> Index;

// ----------------------------------------------------------------------------
// StringIndexer.buildIndexPara (StringIndexer.java:47-83) with
// StringIndexing.run (:30-44). Q1: task i indexes [i*size, (i+1)*size]
// inclusive, so position (i+1)*size is inserted twice; per-key lists are
// concatenated in task order.
// ----------------------------------------------------------------------------
static Index buildIndexPara(const std::string& s, int K, int threadNum) {
  struct Task {
    int from, to;
    Index index;
  };
  std::vector<Task> tasks(threadNum);
  int size = (int)s.size() / threadNum;
  for (int i = 0; i < threadNum - 1; i++) {
    tasks[i].from = i * size;
    tasks[i].to = (i + 1) * size;
  }
  tasks[threadNum - 1].from = (threadNum - 1) * size;
  tasks[threadNum - 1].to = (int)s.size() - K;
  for (auto& t : tasks) {  // StringIndexing.run
    int to = t.to;
    if (to > (int)s.size() - K) to = (int)s.size() - K;
    for (int i = t.from; i <= to; i++) t.index[s.substr(i, K)].push_back(i);
  }
  Index indexTemp = std::move(tasks[0].index);
  for (int n = 1; n < threadNum; n++) {
    for (auto& e : tasks[n].index) {
      auto it = indexTemp.find(e.first);
      if (it == indexTemp.end())
        indexTemp.emplace(e.first, std::move(e.second));
      else
        it->second.insert(it->second.end(), e.second.begin(), e.second.end());
    }
  }
  return indexTemp;
}

// ----------------------------------------------------------------------------
// ReferenceAlignment: Reference(name, aln) (Reference.java:33-59) and
// buildAlnIndex / AlnIndexBuilder.run (ReferenceAlignment.java:104-209):
// k-mer -> sorted distinct alignment columns over all references.
// ----------------------------------------------------------------------------
static Index buildAlnIndex(const std::vector<std::string>& alns, int K) {
  std::unordered_map<std::string, std::set<int>> tmp;
  for (const std::string& aln : alns) {
    std::string seq;
    std::vector<int> seqToAln;
    for (size_t i = 0; i < aln.size(); i++) {
      if (aln[i] != '-') {
        seqToAln.push_back((int)i);
        seq.push_back(aln[i]);
      }
    }
    for (int i = 0; i <= (int)seq.size() - K; i++) {
      std::string s = seq.substr(i, K);
      if (s.find((char)GAP) != std::string::npos) continue;  // :124 (never true)
      tmp[s].insert(seqToAln[i]);
    }
  }
  Index res;
  for (auto& e : tmp) res[e.first] = std::vector<int>(e.second.begin(), e.second.end());
  return res;
}

// ----------------------------------------------------------------------------
// KMerCounterBase
// ----------------------------------------------------------------------------
struct KMerCounter {
  int K = 5;
  std::vector<int> cutoffs;
  float lowCoef = 0.0f;  // Q3: stays 0 in the single-reference Mapper
  float highCoef = 1.25f;

  // KMerCounterBase.mergeHits (KMerCounterBase.java:337-450)
  std::vector<LongHit> mergeHits(const std::vector<const std::vector<int>*>& hits) const {
    static const std::vector<int> EMPTY;
    // LongHit objects live in `pool`; lists hold indices (object identity).
    std::vector<LongHit> pool;
    std::vector<int> longHits, currSet, nextSet;
    size_t nh = hits.size();
    size_t firstMatch = 0;
    do {  // :342-353
      if (firstMatch == nh) return {};
      for (int pos_j : *hits[firstMatch]) {
        pool.push_back({pos_j, (int)firstMatch, 0});
        currSet.push_back((int)pool.size() - 1);
        longHits.push_back((int)pool.size() - 1);
      }
      firstMatch++;
    } while (currSet.empty());
    for (size_t i = firstMatch; i < nh; i++) {  // :354-421
      const std::vector<int>& hits_i = *hits[i];
      if (hits_i.empty()) {
        currSet.clear();
        nextSet.clear();
        continue;
      }
      if (currSet.empty()) {
        for (int pos_j : hits_i) {
          pool.push_back({pos_j, (int)i, 0});
          currSet.push_back((int)pool.size() - 1);
          longHits.push_back((int)pool.size() - 1);
        }
        continue;
      }
      size_t it = 0;  // iteratorSet
      int lHit = currSet[it++];
      size_t j = 0;
      int rHit = hits_i[j];
      while (true) {
        if (pool[lHit].genomePos + pool[lHit].len + 1 < rHit) {  // :375
          if (it < currSet.size()) {
            lHit = currSet[it++];
          } else {
            pool.push_back({rHit, (int)i, 0});
            nextSet.push_back((int)pool.size() - 1);
            longHits.push_back((int)pool.size() - 1);
            j++;
            break;
          }
        } else if (pool[lHit].genomePos + pool[lHit].len + 1 > rHit) {  // :386
          j++;
          pool.push_back({rHit, (int)i, 0});
          nextSet.push_back((int)pool.size() - 1);
          longHits.push_back((int)pool.size() - 1);
          if (j < hits_i.size())
            rHit = hits_i[j];
          else
            break;
        } else {  // :396
          nextSet.push_back(lHit);
          pool[lHit].len++;
          j++;
          if (it < currSet.size())
            lHit = currSet[it++];
          else
            break;
          if (j < hits_i.size())
            rHit = hits_i[j];
          else
            break;
        }
      }
      while (j < hits_i.size()) {  // :412-418
        rHit = hits_i[j];
        pool.push_back({rHit, (int)i, 0});
        nextSet.push_back((int)pool.size() - 1);
        longHits.push_back((int)pool.size() - 1);
        j++;
      }
      currSet.swap(nextSet);
      nextSet.clear();
    }
    // :422 TreeSet ordered by (genomePos, readPos); equal keys keep the first (Q2)
    auto cmp = [&pool](int a, int b) {
      if (pool[a].genomePos == pool[b].genomePos) return pool[a].readPos < pool[b].readPos;
      return pool[a].genomePos < pool[b].genomePos;
    };
    std::set<int, decltype(cmp)> sortedSet(cmp);
    for (int id : longHits) sortedSet.insert(id);
    std::vector<LongHit> res;
    int prev = *sortedSet.begin();
    sortedSet.erase(sortedSet.begin());
    while (!sortedSet.empty()) {  // :425-447
      int curr = *sortedSet.begin();
      sortedSet.erase(sortedSet.begin());
      LongHit& p = pool[prev];
      LongHit& c = pool[curr];
      if (p.genomePos + p.len + K - 1 >= c.genomePos) {
        if (p.len < c.len) {
          p.len = c.genomePos - p.genomePos - K;
          if (p.len >= 0) res.push_back(p);
          prev = curr;
        } else {
          int lenDiff = p.genomePos + p.len + K - c.genomePos;
          c.len -= lenDiff;
          if (c.len >= 0) {
            c.genomePos += lenDiff;
            c.readPos += lenDiff;
            sortedSet.insert(curr);  // silently dropped on key collision
          }
        }
      } else {
        res.push_back(p);
        prev = curr;
      }
    }
    res.push_back(pool[prev]);
    return res;
  }

  // KMerCounterBase.getHits (:452-466)
  std::vector<LongHit> getHits(const std::string& read, const Index& index) const {
    static const std::vector<int> EMPTY;
    int readLen = (int)read.size();
    if (readLen < K) return {};
    std::vector<const std::vector<int>*> matches(readLen - K + 1);
    for (int i = 0; i <= readLen - K; i++) {
      auto it = index.find(read.substr(i, K));
      matches[i] = (it == index.end()) ? &EMPTY : &it->second;
    }
    return mergeHits(matches);
  }

  // KMerCounterBase.concordanceArray (:266-279)
  std::vector<uint8_t> concordanceArray(const std::vector<LongHit>& longHits) const {
    size_t totalHit = longHits.size();
    std::vector<uint8_t> isAdjacent(totalHit, 0);
    const LongHit* prev = &longHits[0];
    for (size_t i = 1; i < totalHit; i++) {
      const LongHit* h_j = &longHits[i];
      volatile float num = (float)(h_j->genomePos - prev->genomePos);
      volatile float den = (float)(h_j->readPos - prev->readPos);
      volatile float rate = num / den;
      if (lowCoef <= rate && rate <= highCoef) isAdjacent[i - 1] = 1;
      prev = h_j;
    }
    return isAdjacent;
  }

  // KMerCounterBase.getInitState(from,to,...) (:13-36)
  Location getInitState(int from, int to, int readLen, const std::vector<LongHit>& longHits,
                        const std::vector<uint8_t>& isAdjacent) const {
    Location res;
    int totalHit = (int)longHits.size();
    const LongHit& h_i = longHits[0];
    res.start = std::max(from, h_i.genomePos - h_i.readPos * 3 / 2);
    res.end = std::min(h_i.genomePos + h_i.len + K + (readLen - h_i.readPos - h_i.len - K) * 3 / 2, to);
    res.startIndex = 0;
    int count = h_i.len;
    int j;
    for (j = 1; j < totalHit; j++) {
      const LongHit& h_j = longHits[j];
      if (h_j.genomePos + K + h_j.len > res.end) break;  // Q4
      count += h_j.len;
      if (isAdjacent[j - 1]) count++;
    }
    res.endIndex = j - 1;
    res.count = count;
    return res;
  }

  // KMerCounterBase.updateCount (:38-158) — the exact incremental state machine (Q5)
  void updateCount(int start, int end, const std::vector<LongHit>& longHits, Location& location,
                   const std::vector<uint8_t>& isAdjacent) const {
    int totalHit = (int)longHits.size();
    int j;
    int count = location.count;
    int sIndex;
    if (start != location.start) {
      sIndex = -1;
      bool prevInWindow;
      const LongHit& startHit = longHits[location.startIndex];
      if (start < location.start) {
        prevInWindow = startHit.genomePos + K + startHit.len <= end;
        for (j = location.startIndex - 1; j >= 0; j--) {
          const LongHit& h_j = longHits[j];
          if (h_j.genomePos < start) {
            sIndex = j + 1;
            break;
          }
          if (h_j.genomePos + K + h_j.len > end) continue;
          if (prevInWindow && isAdjacent[j]) count++;
          count += h_j.len;
          prevInWindow = true;
        }
        if (sIndex == -1) sIndex = 0;
      } else {
        const LongHit* h_j = &longHits[location.startIndex];
        prevInWindow = h_j->genomePos >= start;
        if (prevInWindow) {
          sIndex = location.startIndex;
        } else {
          count -= h_j->len;
          for (j = location.startIndex + 1; j <= location.endIndex; j++) {
            h_j = &longHits[j];
            if (isAdjacent[j - 1]) count--;
            if (h_j->genomePos >= start) break;
            count -= h_j->len;
          }
          if (j <= location.endIndex) {
            sIndex = j;
          } else {
            for (sIndex = j; sIndex < totalHit; sIndex++) {
              h_j = &longHits[sIndex];
              if (h_j->genomePos >= start) break;
            }
          }
        }
      }
      location.start = start;
    } else {
      sIndex = location.startIndex;
    }
    if (end != location.end) {
      int eIndex = -1;
      bool prevInWindow;
      if (end > location.end) {
        const LongHit& endHit = longHits[location.endIndex];
        prevInWindow = endHit.genomePos >= start;
        for (j = location.endIndex + 1; j < totalHit; j++) {
          const LongHit& h_j = longHits[j];
          if (h_j.genomePos + K + h_j.len > end) {
            eIndex = j - 1;
            break;
          }
          if (h_j.genomePos < start) continue;
          if (prevInWindow && isAdjacent[j - 1]) count++;
          count += h_j.len;
          prevInWindow = true;
        }
        if (eIndex == -1) eIndex = totalHit - 1;
      } else {
        const LongHit* h_j = &longHits[location.endIndex];
        if (h_j->genomePos + K + h_j->len > end) {
          count -= h_j->len;
          for (j = location.endIndex - 1; j >= location.startIndex; j--) {
            h_j = &longHits[j];
            if (isAdjacent[j]) count--;
            if (h_j->genomePos + K + h_j->len <= end) break;
            count -= h_j->len;
          }
          if (j >= location.startIndex) {
            eIndex = j;
          } else {
            for (eIndex = j; eIndex >= 0; eIndex--) {
              h_j = &longHits[eIndex];
              if (h_j->genomePos + K + h_j->len <= end) break;
            }
          }
        } else {
          eIndex = location.endIndex;
        }
      }
      location.end = end;
      location.endIndex = eIndex;
    }
    location.startIndex = sIndex;
    location.count = count;
  }

  // KMerCounterBase.computeCount(from,to,...) (:160-166)
  void computeCount(int from, int to, int readLen, const std::vector<LongHit>& longHits, int i,
                    Location& location, const std::vector<uint8_t>& isAdjacent) const {
    const LongHit& h_i = longHits[i];
    int start = std::max(from, h_i.genomePos - h_i.readPos * 3 / 2);
    int end = std::min(h_i.genomePos + h_i.len + K + (readLen - h_i.readPos - h_i.len - K) * 3 / 2, to);
    updateCount(start, end, longHits, location, isAdjacent);
  }

  // KMerCounterBase.computeCount(contigEnds,...) (:169-191)
  void computeCount(const std::vector<int>& contigEnds, int readLen, const std::vector<LongHit>& longHits,
                    int i, Location& location, const std::vector<uint8_t>& isAdjacent) const {
    const LongHit& h_i = longHits[i];
    int start = 0, end = 0;
    for (size_t j = 0, s = 0; j < contigEnds.size(); j++) {
      int e = contigEnds[j];
      if ((int)s <= h_i.genomePos && h_i.genomePos < e) {
        start = std::max((int)s, h_i.genomePos - h_i.readPos * 3 / 2);
        end = std::min(h_i.genomePos + h_i.len + K + (readLen - h_i.readPos - h_i.len - K) * 3 / 2, e);
        if (h_i.genomePos + h_i.len + K > e) return;
        break;
      }
      s = e;
    }
    updateCount(start, end, longHits, location, isAdjacent);
  }

  // KMerCounterBase.pruneWindows (:317-335) + CoordComparator (:253-264)
  static std::vector<MappedRead> pruneWindows(std::vector<MappedRead>& bestPositions) {
    std::stable_sort(bestPositions.begin(), bestPositions.end(), [](const MappedRead& a, const MappedRead& b) {
      if (a.start == b.start) return (b.end - a.end) < 0;
      return (a.start - b.start) < 0;
    });
    std::vector<MappedRead> res;
    MappedRead pivot = bestPositions[0];
    for (size_t i = 1; i < bestPositions.size(); i++) {
      const MappedRead& curr = bestPositions[i];
      if (pivot.end > curr.start) {
        if (pivot.count < curr.count) pivot = curr;
      } else {
        res.push_back(pivot);
        pivot = curr;
      }
    }
    res.push_back(pivot);
    return res;
  }

  // KMerCounter.getInitState(contigEnds,...) (KMerCounter.java:72-116)
  Location getInitStateContigs(const std::vector<int>& contigEnds, int readLen,
                               const std::vector<LongHit>& longHits,
                               const std::vector<uint8_t>& isAdjacent) const {
    Location res;
    int totalHit = (int)longHits.size();
    const LongHit* h_i = nullptr;
    bool done = false;
    for (int i = 0; i < totalHit && !done; i++) {
      h_i = &longHits[i];
      for (size_t j = 0, start = 0; j < contigEnds.size(); j++) {
        int end = contigEnds[j];
        if ((int)start <= h_i->genomePos && h_i->genomePos < end) {
          res.start = std::max((int)start, h_i->genomePos - h_i->readPos * 3 / 2);
          res.end = std::min(h_i->genomePos + h_i->len + K + (readLen - h_i->readPos - h_i->len - K) * 3 / 2, end);
          if (h_i->genomePos + h_i->len + K <= res.end) {
            done = true;  // break outer
            break;
          }
        }
        start = end;
      }
    }
    res.startIndex = 0;
    int count = h_i->len;
    int j;
    for (j = 1; j < totalHit; j++) {
      const LongHit& h_j = longHits[j];
      if (h_j.genomePos + K + h_j.len > res.end) break;
      count += h_j.len;
      if (isAdjacent[j - 1]) count++;
    }
    res.endIndex = j - 1;
    res.count = count;
    return res;
  }

  // KMerCounter.splitLongHits (KMerCounter.java:142-185). The right part of a
  // split is the same object mutated in place.
  std::vector<LongHit> splitLongHits(std::vector<LongHit> longHits, const std::vector<int>& contigEnds) const {
    int contigEnd = contigEnds[0];
    size_t i = 0, j = 0;
    LongHit* hit = &longHits[0];
    std::vector<LongHit> splitted;
    while (true) {
      if (hit->genomePos < contigEnd) {
        if (hit->genomePos + K + hit->len > contigEnd) {
          if (contigEnd - hit->genomePos >= K) {
            LongHit newHit{hit->genomePos, hit->readPos, contigEnd - hit->genomePos - K};
            splitted.push_back(newHit);
          }
          if (hit->genomePos + K + hit->len - contigEnd >= K) {
            hit->readPos += contigEnd - hit->genomePos;
            hit->len -= contigEnd - hit->genomePos;
            hit->genomePos = contigEnd;
          } else {
            j++;
            if (j < longHits.size())
              hit = &longHits[j];
            else
              break;
          }
        } else {
          splitted.push_back(*hit);
          j++;
          if (j < longHits.size())
            hit = &longHits[j];
          else
            break;
        }
      } else {
        i++;
        contigEnd = contigEnds[i];
      }
    }
    return splitted;
  }

  // KMerCounter.moveWindow (KMerCounter.java:187-221)
  bool moveWindow(const std::vector<LongHit>& longHits, const std::vector<uint8_t>& isAdjacent,
                  const std::vector<int>& contigEnds, Location& location, int readLen,
                  std::vector<MappedRead>& out) const {
    int totalHit = (int)longHits.size();
    int cutoff = cutoffs.at(readLen);
    std::vector<MappedRead> bestPositions;
    auto emit = [&]() {
      MappedRead m;
      m.start = location.start;
      m.end = location.end;
      m.count = location.count;
      bestPositions.push_back(m);
    };
    if (location.count > cutoff) emit();
    for (int i = 1; i < totalHit; i++) {
      computeCount(contigEnds, readLen, longHits, i, location, isAdjacent);
      if (location.count > cutoff) emit();
    }
    if (bestPositions.empty()) return false;
    out = pruneWindows(bestPositions);
    return true;
  }

  // KMerCounter.getNBestRegions (KMerCounter.java:223-237); empty vector == null
  std::vector<MappedRead> getNBestRegions(const std::string& read, const Index& index,
                                          const std::vector<int>& contigEnds) const {
    std::vector<LongHit> longHits = getHits(read, index);
    if (longHits.empty()) return {};
    longHits = splitLongHits(longHits, contigEnds);
    if (longHits.empty()) return {};
    std::vector<uint8_t> isAdjacent = concordanceArray(longHits);
    Location location = getInitStateContigs(contigEnds, (int)read.size(), longHits, isAdjacent);
    std::vector<MappedRead> out;
    if (!moveWindow(longHits, isAdjacent, contigEnds, location, (int)read.size(), out)) return {};
    return out;
  }

  // KMerCounterMSA.getNBestRegions (KMerCounterMSA.java:71-97)
  std::vector<MappedRead> getNBestRegionsMSA(const std::string& read, const Index& index, int from, int to,
                                             int minCount) const {
    std::vector<LongHit> hits = getHits(read, index);
    if (hits.empty()) return {};
    int readLen = (int)read.size();
    int totalHit = (int)hits.size();
    std::vector<uint8_t> isConcordant = concordanceArray(hits);
    Location location = getInitState(from, to, readLen, hits, isConcordant);
    std::vector<MappedRead> bestPositions;
    auto emit = [&]() {
      MappedRead m;
      m.start = location.start;
      m.end = location.end;
      m.count = location.count;
      bestPositions.push_back(m);
    };
    if (location.count > minCount) emit();
    for (int i = 1; i < totalHit; i++) {
      computeCount(from, to, readLen, hits, i, location, isConcordant);
      if (location.count > minCount) emit();
    }
    if (bestPositions.empty()) return {};
    return pruneWindows(bestPositions);
  }

  // KMerCounterBase.computeKMerCount (:468-487) — score-only seeding for the random model
  int computeKMerCount(const std::string& read, const Index& index, int from, int to) const {
    int readLen = (int)read.size();
    std::vector<LongHit> matches = getHits(read, index);
    int totalHit = (int)matches.size();
    if (totalHit == 0) return 0;
    std::vector<uint8_t> isAdjacent = concordanceArray(matches);
    Location state = getInitState(from, to, readLen, matches, isAdjacent);
    int maxCount = state.count;
    for (int i = 1; i < totalHit; i++) {
      computeCount(from, to, readLen, matches, i, state, isAdjacent);
      if (state.count > maxCount) maxCount = state.count;
    }
    return maxCount;
  }
};

// ----------------------------------------------------------------------------
// SmithWatermanGotoh (SmithWatermanGotoh.java:51-300)
// ----------------------------------------------------------------------------
struct SmithWatermanGotoh {
  int match = 2, mismatch = -3, gop = 5, gep = 2;
  enum { STOP = 0, LEFT = 1, DIAGONAL = 2, UP = 3 };
  static constexpr int NEG_INF = -(1 << 28);  // stands in for Float.NEGATIVE_INFINITY

  // align (:51-76) = construct (:79-154) + traceback (:157-222)
  Alignment align(const uint8_t* s1, int len1, const uint8_t* s2, int len2,
                  std::vector<uint8_t>* pointersOut = nullptr) const {
    int m = len1 + 1, n = len2 + 1;
    std::vector<uint8_t> pointers((size_t)m * n, STOP);
    std::vector<int> sizesOfVerticalGaps((size_t)m * n, 1), sizesOfHorizontalGaps((size_t)m * n, 1);
    std::vector<int> g(n, NEG_INF), v(n, 0);
    int cellRow = 0, cellCol = 0, cellScore = 0;
    for (int i = 1, k = n; i < m; i++, k += n) {
      int h = NEG_INF;
      int vDiagonal = v[0];
      for (int j = 1, l = k + 1; j < n; j++, l++) {
        int similarityScore = (s1[i - 1] == s2[j - 1]) ? match : mismatch;
        int f = vDiagonal + similarityScore;
        int g1 = g[j] - gep, g2 = v[j] - gop;
        if (g1 > g2) {
          g[j] = g1;
          sizesOfVerticalGaps[l] = sizesOfVerticalGaps[l - n] + 1;
        } else {
          g[j] = g2;
        }
        int h1 = h - gep, h2 = v[j - 1] - gop;
        if (h1 > h2) {
          h = h1;
          sizesOfHorizontalGaps[l] = sizesOfHorizontalGaps[l - 1] + 1;
        } else {
          h = h2;
        }
        vDiagonal = v[j];
        v[j] = std::max(std::max(f, g[j]), std::max(h, 0));
        if (v[j] == 0)
          pointers[l] = STOP;
        else if (v[j] == f)
          pointers[l] = DIAGONAL;
        else if (v[j] == g[j])
          pointers[l] = UP;
        else
          pointers[l] = LEFT;
        if (v[j] > cellScore) {
          cellRow = i;
          cellCol = j;
          cellScore = v[j];
        }
      }
    }
    // traceback (:157-222)
    Alignment alignment;
    alignment.score = cellScore;
    std::string reversed1;
    int identity = 0;
    int i = cellRow;
    alignment.end1 = i;
    int j = cellCol;
    alignment.end2 = j;
    int k = i * n;
    bool stillGoing = true;
    while (stillGoing) {
      switch (pointers[k + j]) {
        case UP:
          for (int l = 0, len = sizesOfVerticalGaps[k + j]; l < len; l++) {
            reversed1.push_back((char)T.lower[s1[--i]]);
            k -= n;
          }
          break;
        case DIAGONAL: {
          uint8_t c1 = s1[--i];
          uint8_t c2 = s2[--j];
          k -= n;
          reversed1.push_back((char)c1);
          if (c1 == c2) identity++;
          break;
        }
        case LEFT:
          for (int l = 0, len = sizesOfHorizontalGaps[k + j]; l < len; l++) {
            reversed1.push_back((char)GAP);
            --j;
          }
          break;
        case STOP:
          stillGoing = false;
      }
    }
    alignment.sequence1.assign(reversed1.rbegin(), reversed1.rend());
    alignment.length = (int)alignment.sequence1.size();
    alignment.start1 = i;
    alignment.start2 = j;
    alignment.identity = identity;
    if (pointersOut) *pointersOut = pointers;
    return alignment;
  }

  // findMaxScore (:240-300)
  int findMaxScore(const uint8_t* s1, int len1, const uint8_t* s2, int len2) const {
    int m = len1 + 1, n = len2 + 1;
    std::vector<int> g(n, NEG_INF), v(n, 0);
    int maxScore = 0;
    for (int i = 1; i < m; i++) {
      int h = NEG_INF;
      int vDiagonal = v[0];
      for (int j = 1; j < n; j++) {
        int similarityScore = (s1[i - 1] == s2[j - 1]) ? match : mismatch;
        int f = vDiagonal + similarityScore;
        int g1 = g[j] - gep, g2 = v[j] - gop;
        g[j] = (g1 > g2) ? g1 : g2;
        int h1 = h - gep, h2 = v[j - 1] - gop;
        h = (h1 > h2) ? h1 : h2;
        vDiagonal = v[j];
        v[j] = std::max(std::max(f, g[j]), std::max(h, 0));
        if (v[j] > maxScore) maxScore = v[j];
      }
    }
    return maxScore;
  }
};

// ----------------------------------------------------------------------------
// Reference (Reference.java) — the fields the hot path reads
// ----------------------------------------------------------------------------
struct Reference {
  std::string seq;
  std::vector<int> contigEnds, fragmentEnds;
  bool isFragmented = false;
  Index index;
};

enum PairClass { CLS_CONC0 = 0, CLS_CONC1 = 1, CLS_DISCORDANT = 2, CLS_LEFT = 3, CLS_RIGHT = 4, CLS_UNMAPPED = 5 };

struct PairResult {
  int cls = CLS_UNMAPPED;
  bool has1 = false, has2 = false;
  MappedRead r1, r2;
};

// ----------------------------------------------------------------------------
// Mapper.ReadsMapping.mapRead (Mapper.java:92-390)
// ----------------------------------------------------------------------------
struct Mapper {
  KMerCounter counter;
  SmithWatermanGotoh aligner;
  int insertLen = 1000;

  std::vector<MappedRead> mapReadFast(const std::string& read, const Reference& genome) const {  // :56-65
    return counter.getNBestRegions(read, genome.index, genome.contigEnds);
  }

  void alignTo(MappedRead& r, const std::string& seq, const Reference& genome) const {
    // Arrays.copyOfRange(genome.seqB, start, end) + aligner.align (:236-239)
    std::string sub(r.end - r.start, '\0');
    for (int p = r.start; p < r.end; p++)
      if (p < (int)genome.seq.size()) sub[p - r.start] = genome.seq[p];
    r.aln = aligner.align((const uint8_t*)seq.data(), (int)seq.size(), (const uint8_t*)sub.data(), (int)sub.size());
    r.hasAln = true;
  }

  // The decision of mapRead (Mapper.java:100-389) on the four candidate lists, without the alignments: which class, and
  // which candidate (side 0 = forward list, 1 = reverse list; index) becomes r1 / r2 (-1 = null). chooseFragment
  // (MappedRead.java:39-50) is applied to the lists in place, as in the Java.
  struct Decision {
    int cls = CLS_UNMAPPED, sd1 = -1, i1 = -1, sd2 = -1, i2 = -1;
  };
  Decision decide(std::vector<MappedRead>& forward1, std::vector<MappedRead>& reverse1, std::vector<MappedRead>& forward2,
                  std::vector<MappedRead>& reverse2, const Reference& genome) const {
    int maxCount = 0, max1 = -1, max2 = -1, side = -1;
    int maxInd1 = -1, count1 = 0, side1 = -1;
    int maxInd2 = -1, count2 = 0, side2 = -1;
    std::vector<int> maxIndFrag1, countFrag1, sideFrag1, maxIndFrag2, countFrag2, sideFrag2;
    size_t nf = genome.fragmentEnds.size();
    if (genome.isFragmented) {
      maxIndFrag1.assign(nf, 0);
      countFrag1.assign(nf, 0);
      sideFrag1.assign(nf, 0);
      maxIndFrag2.assign(nf, 0);
      countFrag2.assign(nf, 0);
      sideFrag2.assign(nf, 0);
    }
    auto scan1 = [&](std::vector<MappedRead>& v, int sd) {  // :124-157
      for (size_t j = 0; j < v.size(); j++) {
        MappedRead& read1 = v[j];
        if (genome.isFragmented) {
          read1.chooseFragment(genome.fragmentEnds);
          int f = read1.fragmentID;
          if (read1.count > countFrag1[f]) {
            countFrag1[f] = read1.count;
            sideFrag1[f] = sd;
            maxIndFrag1[f] = (int)j;
          }
        }
        if (read1.count > count1) {
          count1 = read1.count;
          side1 = sd;
          maxInd1 = (int)j;
        }
      }
    };
    auto scan2 = [&](std::vector<MappedRead>& v, int sd) {  // :158-191
      for (size_t k = 0; k < v.size(); k++) {
        MappedRead& read2 = v[k];
        if (genome.isFragmented) {
          read2.chooseFragment(genome.fragmentEnds);
          int f = read2.fragmentID;
          if (read2.count > countFrag2[f]) {
            countFrag2[f] = read2.count;
            sideFrag2[f] = sd;
            maxIndFrag2[f] = (int)k;
          }
        }
        if (read2.count > count2) {
          count2 = read2.count;
          side2 = sd;
          maxInd2 = (int)k;
        }
      }
    };
    scan1(forward1, 0);
    scan1(reverse1, 1);
    scan2(reverse2, 1);  // mate 2: reverse first, then forward
    scan2(forward2, 0);
    for (size_t j = 0; j < forward1.size(); j++) {  // :193-208
      const MappedRead& read1 = forward1[j];
      for (size_t k = 0; k < reverse2.size(); k++) {
        const MappedRead& read2 = reverse2[k];
        if (read1.start < read2.end && read2.end - read1.start <= insertLen) {
          if (read1.fragmentID == read2.fragmentID && read1.count + read2.count > maxCount) {
            maxCount = read1.count + read2.count;
            max1 = (int)j;
            max2 = (int)k;
            side = 0;
          }
        }
      }
    }
    for (size_t j = 0; j < reverse1.size(); j++) {  // :209-224
      const MappedRead& read1 = reverse1[j];
      for (size_t k = 0; k < forward2.size(); k++) {
        const MappedRead& read2 = forward2[k];
        if (read2.start < read1.end && read1.end - read2.start <= insertLen) {
          if (read1.fragmentID == read2.fragmentID && read1.count + read2.count > maxCount) {
            maxCount = read1.count + read2.count;
            max1 = (int)j;
            max2 = (int)k;
            side = 1;
          }
        }
      }
    }
    Decision d;
    switch (side) {
      case 0:  // :226-244
        d.cls = CLS_CONC0, d.sd1 = 0, d.i1 = max1, d.sd2 = 1, d.i2 = max2;
        break;
      case 1:  // :245-263
        d.cls = CLS_CONC1, d.sd1 = 1, d.i1 = max1, d.sd2 = 0, d.i2 = max2;
        break;
      default:
        if (side1 == -1) {
          if (side2 == -1) {
            d.cls = CLS_UNMAPPED;  // :267
          } else {
            d.cls = CLS_RIGHT;  // :271-287
            d.sd2 = side2, d.i2 = maxInd2;
          }
        } else {
          if (side2 == -1) {
            d.cls = CLS_LEFT;  // :291-307
            d.sd1 = side1, d.i1 = maxInd1;
          } else if (genome.isFragmented) {  // :309-353
            int maxSum = 0, maxIndex = -1;
            for (size_t i = 0; i < nf; i++) {
              if (countFrag1[i] > 0 && countFrag2[i] > 0 && countFrag1[i] + countFrag2[i] > maxSum) {
                maxSum = countFrag1[i] + countFrag2[i];
                maxIndex = (int)i;
              }
            }
            if (maxIndex != -1) {
              d.cls = CLS_DISCORDANT;
              d.sd1 = sideFrag1[maxIndex], d.i1 = maxIndFrag1[maxIndex];
              d.sd2 = sideFrag2[maxIndex], d.i2 = maxIndFrag2[maxIndex];
            } else {
              d.cls = CLS_UNMAPPED;
            }
          } else {  // :355-385
            d.cls = CLS_DISCORDANT;
            d.sd1 = side1, d.i1 = maxInd1;
            d.sd2 = side2, d.i2 = maxInd2;
          }
        }
    }
    return d;
  }

  PairResult mapRead(const std::string& seq1, const std::string& seq2, const Reference& genome) const {
    std::string rc1 = getComplement(seq1), rc2 = getComplement(seq2);
    std::vector<MappedRead> forward1 = mapReadFast(seq1, genome);
    std::vector<MappedRead> reverse1 = mapReadFast(rc1, genome);
    std::vector<MappedRead> forward2 = mapReadFast(seq2, genome);
    std::vector<MappedRead> reverse2 = mapReadFast(rc2, genome);
    const Decision d = decide(forward1, reverse1, forward2, reverse2, genome);
    PairResult res;
    res.cls = d.cls;
    if (d.i1 >= 0) {
      res.r1 = (d.sd1 == 0) ? forward1[d.i1] : reverse1[d.i1];
      res.r1.reverse = d.sd1;
      res.has1 = true;
      alignTo(res.r1, d.sd1 == 0 ? seq1 : rc1, genome);
    }
    if (d.i2 >= 0) {
      res.r2 = (d.sd2 == 0) ? forward2[d.i2] : reverse2[d.i2];
      res.r2.reverse = d.sd2;
      res.has2 = true;
      alignTo(res.r2, d.sd2 == 0 ? seq2 : rc2, genome);
    }
    return res;
  }
};

// ----------------------------------------------------------------------------
// MapperMSA.ReadMapping.mapRead (MapperMSA.java:56-244)
// ----------------------------------------------------------------------------
struct MapperMSA {
  KMerCounter counterRF;
  int insertLen = 1000;

  std::vector<MappedRead> mapReadFast(const std::string& read, const Index& index, int msaLength) const {
    if (read == "*") return {};  // :57 Q6
    int readLen = (int)read.size();
    return counterRF.getNBestRegionsMSA(read, index, 0, msaLength, counterRF.cutoffs.at(readLen));
  }

  PairResult mapRead(const std::string& seq1, const std::string& seq2, const Index& index, int msaLength) const {
    std::string rc1 = getComplement(seq1), rc2 = getComplement(seq2);
    std::vector<MappedRead> forward1 = mapReadFast(seq1, index, msaLength);
    std::vector<MappedRead> reverse1 = mapReadFast(rc1, index, msaLength);
    std::vector<MappedRead> forward2 = mapReadFast(seq2, index, msaLength);
    std::vector<MappedRead> reverse2 = mapReadFast(rc2, index, msaLength);
    int maxCount = 0, max1 = -1, max2 = -1, side = -1;
    int maxInd1 = -1, count1 = 0, side1 = -1;
    int maxInd2 = -1, count2 = 0, side2 = -1;
    for (size_t j = 0; j < forward1.size(); j++)
      if (forward1[j].count > count1) count1 = forward1[j].count, side1 = 0, maxInd1 = (int)j;
    for (size_t j = 0; j < reverse1.size(); j++)
      if (reverse1[j].count > count1) count1 = reverse1[j].count, side1 = 1, maxInd1 = (int)j;
    for (size_t k = 0; k < reverse2.size(); k++)
      if (reverse2[k].count > count2) count2 = reverse2[k].count, side2 = 1, maxInd2 = (int)k;
    for (size_t k = 0; k < forward2.size(); k++)
      if (forward2[k].count > count2) count2 = forward2[k].count, side2 = 0, maxInd2 = (int)k;
    for (size_t j = 0; j < forward1.size(); j++)
      for (size_t k = 0; k < reverse2.size(); k++) {
        const MappedRead &read1 = forward1[j], &read2 = reverse2[k];
        if (read1.start < read2.end && read2.end - read1.start <= insertLen)
          if (read1.count + read2.count > maxCount) maxCount = read1.count + read2.count, max1 = (int)j, max2 = (int)k, side = 0;
      }
    for (size_t j = 0; j < reverse1.size(); j++)
      for (size_t k = 0; k < forward2.size(); k++) {
        const MappedRead &read1 = reverse1[j], &read2 = forward2[k];
        if (read2.start < read1.end && read1.end - read2.start <= insertLen)
          if (read1.count + read2.count > maxCount) maxCount = read1.count + read2.count, max1 = (int)j, max2 = (int)k, side = 1;
      }
    PairResult res;
    auto take1 = [&](int sd, int ind) {
      res.r1 = (sd == 0) ? forward1[ind] : reverse1[ind];
      res.r1.reverse = sd;
      res.has1 = true;
    };
    auto take2 = [&](int sd, int ind) {
      res.r2 = (sd == 0) ? forward2[ind] : reverse2[ind];
      res.r2.reverse = sd;
      res.has2 = true;
    };
    if (side == 0) {
      take1(0, max1), take2(1, max2), res.cls = CLS_CONC0;
    } else if (side == 1) {
      take1(1, max1), take2(0, max2), res.cls = CLS_CONC1;
    } else if (side1 == -1) {
      if (side2 == -1)
        res.cls = CLS_UNMAPPED;
      else
        res.cls = CLS_RIGHT, take2(side2, maxInd2);
    } else if (side2 == -1) {
      res.cls = CLS_LEFT, take1(side1, maxInd1);
    } else {
      res.cls = CLS_DISCORDANT, take1(side1, maxInd1), take2(side2, maxInd2);
    }
    return res;
  }
};

// ----------------------------------------------------------------------------
// Random model -> cutoffs (harness only; the reference RNG is unseeded, Q10).
// MarkovModel (MarkovModel.java:18-119), RandomModel.genModel (RandomModel.java:83-145),
// KMerCounterBase.computeCuttofs (KMerCounterBase.java:229-251).
// A seeded splitmix64 replaces java.util.Random.
// ----------------------------------------------------------------------------
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  int nextInt(int n) { return (int)(next() % (uint64_t)n); }
  float nextFloat() { return (float)(next() >> 40) / (float)(1 << 24); }
};

struct MarkovModel {
  int K;
  std::unordered_map<std::string, std::array<float, 4>> freqs;
  std::vector<std::string> genomes;
  MarkovModel(const std::vector<std::string>& gs, int order) : K(order), genomes(gs) {
    std::unordered_map<std::string, std::array<int, 4>> counts;
    for (const std::string& genome : genomes) {
      for (int j = 0; j < (int)genome.size() - K; j++) {
        std::string s = genome.substr(j, K);
        bool ok = true;
        for (char c : s)
          if (c != 'A' && c != 'T' && c != 'G' && c != 'C') ok = false;
        if (!ok) continue;
        auto& cnt = counts.emplace(s, std::array<int, 4>{0, 0, 0, 0}).first->second;
        switch (genome[j + K]) {
          case 'A': cnt[0]++; break;
          case 'T': cnt[1]++; break;
          case 'G': cnt[2]++; break;
          case 'C': cnt[3]++; break;
          default: break;
        }
      }
    }
    for (auto& e : counts) {
      int total = e.second[0] + e.second[1] + e.second[2] + e.second[3];
      float cf = 1.0f / total;
      std::array<float, 4> fr;
      fr[0] = e.second[0] * cf;
      fr[1] = fr[0] + e.second[1] * cf;
      fr[2] = fr[1] + e.second[2] * cf;
      fr[3] = 1.0f;
      freqs[e.first] = fr;
    }
  }
  std::string generate(int len, Rng& rnd) const {
    std::string res;
    const std::string& genome = genomes[rnd.nextInt((int)genomes.size())];
    int genomePos = rnd.nextInt((int)genome.size() - K + 1);
    res.append(genome, genomePos, K);
    for (int j = K; j < len; j++) {
      auto it = freqs.find(res.substr(j - K, K));
      if (it == freqs.end()) {
        res.push_back((char)T.iToN[rnd.nextInt(4)]);
      } else {
        float pr = rnd.nextFloat();
        const auto& fr = it->second;
        res.push_back(pr <= fr[0] ? 'A' : pr <= fr[1] ? 'T' : pr <= fr[2] ? 'G' : 'C');
      }
    }
    return res;
  }
};

}  // namespace vo

// ============================================================================
// C interface (ctypes)
// ============================================================================
using namespace vo;

struct OracleCtx {
  Reference genome;
  Mapper mapper;
  MapperMSA mapperMSA;
  Index msaIndex;
  int msaLength = 0;
};

// Layout shared with include/virgena_gpu.h (vg_mate_t / vg_pair_result_t).
struct MateRec {
  int32_t present, start, end, count, reverse, fragment_id;
  int32_t score, start1, end1, start2, end2, identity, length, pad_;
  int64_t seq1_offset;
};
struct PairRec {
  int32_t cls, pad_;
  MateRec m[2];
};

static void fillMate(MateRec& o, bool has, const MappedRead& r) {
  memset(&o, 0, sizeof o);
  o.seq1_offset = -1;
  if (!has) return;
  o.present = 1;
  o.start = r.start;
  o.end = r.end;
  o.count = r.count;
  o.reverse = r.reverse;
  o.fragment_id = r.fragmentID;
  if (r.hasAln) {
    o.score = r.aln.score;
    o.start1 = r.aln.start1;
    o.end1 = r.aln.end1;
    o.start2 = r.aln.start2;
    o.end2 = r.aln.end2;
    o.identity = r.aln.identity;
    o.length = r.aln.length;
  }
}

extern "C" {

OracleCtx* vo_create(int K, float lowCoef, float highCoef, const int32_t* cutoffs, int nCutoffs, int insertLen,
                     int match, int mismatch, int gop, int gep) {
  OracleCtx* c = new OracleCtx();
  c->mapper.counter.K = K;
  c->mapper.counter.lowCoef = lowCoef;
  c->mapper.counter.highCoef = highCoef;
  c->mapper.counter.cutoffs.assign(cutoffs, cutoffs + nCutoffs);
  c->mapper.insertLen = insertLen;
  c->mapper.aligner.match = match;
  c->mapper.aligner.mismatch = mismatch;
  c->mapper.aligner.gop = gop;
  c->mapper.aligner.gep = gep;
  c->mapperMSA.counterRF = c->mapper.counter;
  c->mapperMSA.insertLen = insertLen;
  return c;
}

void vo_destroy(OracleCtx* c) { delete c; }

// Reference(...) + StringIndexer.buildIndexPara
void vo_set_reference(OracleCtx* c, const uint8_t* seq, int G, const int32_t* contigEnds, int nC,
                      const int32_t* fragmentEnds, int nF, int isFragmented, int indexThreadNum) {
  c->genome.seq.assign((const char*)seq, G);
  c->genome.contigEnds.assign(contigEnds, contigEnds + nC);
  c->genome.fragmentEnds.assign(fragmentEnds, fragmentEnds + nF);
  c->genome.isFragmented = isFragmented != 0;
  c->genome.index = buildIndexPara(c->genome.seq, c->mapper.counter.K, indexThreadNum);
}

// Export the k-mer index as CSR sorted by k-mer bytes: keys[nKeys*K], offsets[nKeys+1], positions.
int vo_index_num_keys(OracleCtx* c, int msa) { return (int)(msa ? c->msaIndex : c->genome.index).size(); }
int64_t vo_index_num_positions(OracleCtx* c, int msa) {
  int64_t n = 0;
  for (auto& e : (msa ? c->msaIndex : c->genome.index)) n += (int64_t)e.second.size();
  return n;
}
void vo_index_export(OracleCtx* c, int msa, uint8_t* keys, int64_t* offsets, int32_t* positions) {
  const Index& idx = msa ? c->msaIndex : c->genome.index;
  std::map<std::string, const std::vector<int>*> sorted;
  for (auto& e : idx) sorted[e.first] = &e.second;
  int K = msa ? c->mapperMSA.counterRF.K : c->mapper.counter.K;
  int64_t off = 0;
  int k = 0;
  for (auto& e : sorted) {
    memcpy(keys + (size_t)k * K, e.first.data(), K);
    offsets[k] = off;
    for (int p : *e.second) positions[off++] = p;
    k++;
  }
  offsets[k] = off;
}

// ReferenceAlignment.buildAlnIndex over `nRefs` gapped rows of equal length.
void vo_set_msa(OracleCtx* c, const uint8_t* rows, int nRefs, int alnLength, int K, float lowCoef, float highCoef,
                const int32_t* cutoffs, int nCutoffs) {
  std::vector<std::string> alns(nRefs);
  for (int r = 0; r < nRefs; r++) alns[r].assign((const char*)rows + (size_t)r * alnLength, alnLength);
  c->mapperMSA.counterRF.K = K;
  c->mapperMSA.counterRF.lowCoef = lowCoef;
  c->mapperMSA.counterRF.highCoef = highCoef;
  c->mapperMSA.counterRF.cutoffs.assign(cutoffs, cutoffs + nCutoffs);
  c->msaIndex = buildAlnIndex(alns, K);
  c->msaLength = alnLength;
}

// KMerCounterBase.getHits -> LongHit triples (genomePos, readPos, len); returns count (or needed size).
int vo_get_hits(OracleCtx* c, int msa, const uint8_t* read, int L, int32_t* out, int cap, int doSplit) {
  const KMerCounter& kc = msa ? c->mapperMSA.counterRF : c->mapper.counter;
  std::vector<LongHit> h = kc.getHits(std::string((const char*)read, L), msa ? c->msaIndex : c->genome.index);
  if (doSplit && !h.empty()) h = kc.splitLongHits(h, c->genome.contigEnds);
  for (size_t i = 0; i < h.size() && (int)i < cap; i++) {
    out[3 * i] = h[i].genomePos;
    out[3 * i + 1] = h[i].readPos;
    out[3 * i + 2] = h[i].len;
  }
  return (int)h.size();
}

// getNBestRegions for one read -> (start,end,count) triples.
int vo_seed(OracleCtx* c, int msa, const uint8_t* read, int L, int32_t* out, int cap) {
  std::string r((const char*)read, L);
  std::vector<MappedRead> v =
      msa ? c->mapperMSA.mapReadFast(r, c->msaIndex, c->msaLength) : c->mapper.mapReadFast(r, c->genome);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) {
    out[3 * i] = v[i].start;
    out[3 * i + 1] = v[i].end;
    out[3 * i + 2] = v[i].count;
  }
  return (int)v.size();
}

// KMerCounter.mapReadToRegion (KMerCounter.java:119-140) with from = 0, to = reference / MSA length: the first window of
// maximal count over all hits (no cutoff, no splitLongHits, no pruning). Returns 0 when the read has no hit (null).
// out = {start, end, count}; count equals KMerCounterBase.computeKMerCount (:468-487).
int vo_map_read_to_region(OracleCtx* c, int msa, const uint8_t* read, int L, int32_t* out) {
  const KMerCounter& kc = msa ? c->mapperMSA.counterRF : c->mapper.counter;
  const int to = msa ? c->msaLength : (int)c->genome.seq.size();
  std::string r((const char*)read, L);
  std::vector<LongHit> h = kc.getHits(r, msa ? c->msaIndex : c->genome.index);
  if (h.empty()) return 0;
  std::vector<uint8_t> adj = kc.concordanceArray(h);
  Location st = kc.getInitState(0, to, L, h, adj);
  out[0] = st.start, out[1] = st.end, out[2] = st.count;
  for (int i = 1; i < (int)h.size(); i++) {
    kc.computeCount(0, to, L, h, i, st, adj);
    if (st.count > out[2]) out[0] = st.start, out[1] = st.end, out[2] = st.count;
  }
  if (out[2] != kc.computeKMerCount(r, msa ? c->msaIndex : c->genome.index, 0, to)) return -1;
  return 1;
}

// Walk the windows of one read with the exact incremental state machine and,
// beside it, a direct recount: per hit i writes (start,end,count_incremental,count_direct).
int vo_window_walk(OracleCtx* c, int msa, const uint8_t* read, int L, int32_t* out, int cap) {
  const KMerCounter& kc = msa ? c->mapperMSA.counterRF : c->mapper.counter;
  int K = kc.K;
  std::string r((const char*)read, L);
  std::vector<LongHit> h = kc.getHits(r, msa ? c->msaIndex : c->genome.index);
  if (h.empty()) return 0;
  if (!msa) h = kc.splitLongHits(h, c->genome.contigEnds);
  if (h.empty()) return 0;
  std::vector<uint8_t> adj = kc.concordanceArray(h);
  Location loc = msa ? kc.getInitState(0, c->msaLength, L, h, adj) : kc.getInitStateContigs(c->genome.contigEnds, L, h, adj);
  for (int i = 0; i < (int)h.size(); i++) {
    if (i > 0) {
      if (msa)
        kc.computeCount(0, c->msaLength, L, h, i, loc, adj);
      else
        kc.computeCount(c->genome.contigEnds, L, h, i, loc, adj);
    }
    int direct = 0, prev = -1;
    for (int j = 0; j < (int)h.size(); j++) {
      if (h[j].genomePos >= loc.start && h[j].genomePos + K + h[j].len <= loc.end) {
        direct += h[j].len;
        if (prev == j - 1 && prev >= 0 && adj[j - 1]) direct++;
        prev = j;
      }
    }
    if (i < cap) {
      out[4 * i] = loc.start;
      out[4 * i + 1] = loc.end;
      out[4 * i + 2] = loc.count;
      out[4 * i + 3] = direct;
    }
  }
  return (int)h.size();
}

// KAT hook: concordanceArray + getInitState(from, to) + computeCount(from, to) (KMerCounterBase.java:13-36, 38-166, 266-279)
// over a CRAFTED LongHit list (hits = n x {genomePos, readPos, len}; it need not satisfy the invariants of mergeHits'
// output), so that the literal state machine can be pinned by hand-derived cases (Q4, Q5). Per hit i writes
// {start, end, count, startIndex, endIndex} after the i-th step.
void vo_walk_crafted(int K, float lowCoef, float highCoef, int from, int to, int readLen, const int32_t* hits, int n,
                     int32_t* out) {
  KMerCounter kc;
  kc.K = K, kc.lowCoef = lowCoef, kc.highCoef = highCoef;
  std::vector<LongHit> h(n);
  for (int i = 0; i < n; i++) h[i] = LongHit{hits[3 * i], hits[3 * i + 1], hits[3 * i + 2]};
  if (n == 0) return;
  std::vector<uint8_t> adj = kc.concordanceArray(h);
  Location loc = kc.getInitState(from, to, readLen, h, adj);
  for (int i = 0; i < n; i++) {
    if (i > 0) kc.computeCount(from, to, readLen, h, i, loc, adj);
    out[5 * i] = loc.start, out[5 * i + 1] = loc.end, out[5 * i + 2] = loc.count;
    out[5 * i + 3] = loc.startIndex, out[5 * i + 4] = loc.endIndex;
  }
}

// KAT hook: the decision part of Mapper.mapRead (Mapper.java:100-389) on CRAFTED candidate lists (each n x {start, end,
// count}): out = {class, side of r1 (0 forward1, 1 reverse1, -1 null), index, side of r2, index, fragmentID of r1,
// fragmentID of r2}.
void vo_pair_select(const int32_t* fragmentEnds, int nF, int isFragmented, int insertLen, const int32_t* f1, int nf1,
                    const int32_t* r1, int nr1, const int32_t* f2, int nf2, const int32_t* r2, int nr2, int32_t* out) {
  Reference g;
  g.fragmentEnds.assign(fragmentEnds, fragmentEnds + nF);
  g.isFragmented = isFragmented != 0;
  Mapper m;
  m.insertLen = insertLen;
  auto mk = [](const int32_t* p, int n) {
    std::vector<MappedRead> v(n);
    for (int i = 0; i < n; i++) v[i].start = p[3 * i], v[i].end = p[3 * i + 1], v[i].count = p[3 * i + 2];
    return v;
  };
  std::vector<MappedRead> F1 = mk(f1, nf1), R1 = mk(r1, nr1), F2 = mk(f2, nf2), R2 = mk(r2, nr2);
  const Mapper::Decision d = m.decide(F1, R1, F2, R2, g);
  out[0] = d.cls, out[1] = d.sd1, out[2] = d.i1, out[3] = d.sd2, out[4] = d.i2;
  out[5] = d.i1 >= 0 ? (d.sd1 == 0 ? F1 : R1)[d.i1].fragmentID : -1;
  out[6] = d.i2 >= 0 ? (d.sd2 == 0 ? F2 : R2)[d.i2].fragmentID : -1;
}

// SmithWatermanGotoh.align. seq1 receives `length` bytes; returns length. rec = 7 ints.
int vo_sw_align(int match, int mismatch, int gop, int gep, const uint8_t* s1, int len1, const uint8_t* s2, int len2,
                int32_t* rec, uint8_t* seq1, int cap) {
  SmithWatermanGotoh sw;
  sw.match = match, sw.mismatch = mismatch, sw.gop = gop, sw.gep = gep;
  Alignment a = sw.align(s1, len1, s2, len2);
  rec[0] = a.score, rec[1] = a.start1, rec[2] = a.end1, rec[3] = a.start2, rec[4] = a.end2, rec[5] = a.identity,
  rec[6] = a.length;
  if (a.length <= cap) memcpy(seq1, a.sequence1.data(), a.length);
  return a.length;
}

int vo_sw_max_score(int match, int mismatch, int gop, int gep, const uint8_t* s1, int len1, const uint8_t* s2, int len2) {
  SmithWatermanGotoh sw;
  sw.match = match, sw.mismatch = mismatch, sw.gop = gop, sw.gep = gep;
  return sw.findMaxScore(s1, len1, s2, len2);
}

// Batched SW over independent (read, window) tasks on nThreads host threads (CPU baseline for GCUPS).
void vo_sw_batch(int match, int mismatch, int gop, int gep, const uint8_t* reads, const int64_t* roff,
                 const uint8_t* wins, const int64_t* woff, int n, int32_t* recs, uint8_t* seq1pool,
                 const int64_t* soff, int nThreads) {
  SmithWatermanGotoh sw;
  sw.match = match, sw.mismatch = mismatch, sw.gop = gop, sw.gep = gep;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= n) break;
      Alignment a = sw.align(reads + roff[i], (int)(roff[i + 1] - roff[i]), wins + woff[i], (int)(woff[i + 1] - woff[i]));
      int32_t* rec = recs + 7 * (size_t)i;
      rec[0] = a.score, rec[1] = a.start1, rec[2] = a.end1, rec[3] = a.start2, rec[4] = a.end2, rec[5] = a.identity,
      rec[6] = a.length;
      if (seq1pool) memcpy(seq1pool + soff[i], a.sequence1.data(), a.length);
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nThreads; t++) th.emplace_back(work);
  for (auto& t : th) t.join();
}

// Mapper.mapReads / MapperMSA.mapAllReads over N pairs, batchSize pairs per task on nThreads threads
// (Mapper.java:436-445). Results are written per pair in input order; sequence1 bytes go to seq1pool at
// deterministic offsets (prefix sum of lengths, mate 1 then mate 2). Returns total seq1 bytes (or -needed).
// stats[7] = totalPairs, bothExist, readsTotal, totalMapped, mappedForward, mappedReverse, bothMapped;
// stats64[2] = totalScore, totalConcordant  (Mapper.java:446-473).
int64_t vo_map_pairs(OracleCtx* c, int msa, const uint8_t* bases, const int64_t* off1, const int64_t* off2,
                     const int32_t* len1, const int32_t* len2, int64_t N, int batchSize, int nThreads, PairRec* out,
                     uint8_t* seq1pool, int64_t poolCap, int64_t* stats) {
  std::vector<PairResult> results(N);
  std::atomic<int64_t> next(0);
  int64_t nTasks = (N + batchSize - 1) / batchSize;
  auto work = [&]() {
    for (;;) {
      int64_t t = next.fetch_add(1);
      if (t >= nTasks) break;
      int64_t from = t * batchSize, to = std::min<int64_t>(from + batchSize, N);
      for (int64_t i = from; i < to; i++) {
        std::string s1((const char*)bases + off1[i], len1[i]), s2((const char*)bases + off2[i], len2[i]);
        results[i] = msa ? c->mapperMSA.mapRead(s1, s2, c->msaIndex, c->msaLength) : c->mapper.mapRead(s1, s2, c->genome);
      }
    }
  };
  if (c->genome.seq.empty() && !msa) {
    // Mapper.java:422-426: empty reference -> everything unmapped
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  int64_t pool = 0;
  int64_t totalPairs = N, bothExist = 0, readsTotal = 0, totalMapped = 0, mappedForward = 0, mappedReverse = 0,
          bothMapped = 0, totalScore = 0, totalConcordant = 0;
  for (int64_t i = 0; i < N; i++) {
    const PairResult& r = results[i];
    bool e1 = !(len1[i] == 1 && bases[off1[i]] == '*'), e2 = !(len2[i] == 1 && bases[off2[i]] == '*');
    if (e1 && e2) bothExist++, readsTotal += 2;
    else if (e1 || e2) readsTotal++;
    out[i].cls = r.cls;
    out[i].pad_ = 0;
    fillMate(out[i].m[0], r.has1, r.r1);
    fillMate(out[i].m[1], r.has2, r.r2);
    for (int k = 0; k < 2; k++) {
      bool has = k == 0 ? r.has1 : r.has2;
      const MappedRead& mr = k == 0 ? r.r1 : r.r2;
      if (!has) continue;
      totalMapped++;
      if (mr.hasAln) {
        out[i].m[k].seq1_offset = pool;
        if (seq1pool && pool + mr.aln.length <= poolCap) memcpy(seq1pool + pool, mr.aln.sequence1.data(), mr.aln.length);
        pool += mr.aln.length;
        totalScore += mr.aln.score;
      } else {
        totalScore += mr.count;  // MapperMSA.java:161
      }
    }
    bool conc = r.cls == CLS_CONC0 || r.cls == CLS_CONC1;
    if (conc) totalConcordant++, mappedForward++, mappedReverse++;
    else {
      if (r.has1) (r.r1.reverse ? mappedReverse : mappedForward)++;
      if (r.has2) (r.r2.reverse ? mappedReverse : mappedForward)++;
    }
    if (conc || r.cls == CLS_DISCORDANT) bothMapped++;
  }
  if (stats) {
    stats[0] = totalPairs, stats[1] = bothExist, stats[2] = readsTotal, stats[3] = totalMapped, stats[4] = mappedForward,
    stats[5] = mappedReverse, stats[6] = bothMapped, stats[7] = totalScore, stats[8] = totalConcordant;
  }
  return pool;
}

// ConsensusBuilderSimple.ReadCounting.run (:56-75) + sum (:138-145): counts[G*5] int32, accumulated.
// Returns 0, or -1 if a byte in ('T','a') would have thrown in Java.
int vo_pileup(const PairRec* recs, int64_t N, const uint8_t* seq1pool, int G, int32_t* counts) {
  for (int64_t i = 0; i < N; i++)
    for (int k = 0; k < 2; k++) {
      const MateRec& m = recs[i].m[k];
      if (!m.present) continue;
      const uint8_t* seq1 = seq1pool + m.seq1_offset;
      int j = m.start + m.start2;
      for (int t = 0; t < m.length; t++) {
        uint8_t ch = seq1[t];
        if (ch < 'a') {
          if (ch > 'T' || j >= G) return -1;
          counts[(size_t)j * 5 + T.nToI[ch]]++;
          j++;
        }
      }
    }
  return 0;
}

// ProblemIntervalBuilder.IdentityCounting.run (ProblemIntervalBuilder.java:35-133) + countIdentity (:135-160): sliding-window
// identity of every mapped read summed per window start, in binary32 and in the reference's order: thread t sums its
// reads [t*size, (t+1)*size) (the last one up to the end) in list order into its own array, the arrays are then added in
// thread order. mappedReads = r1, r2 of every pair in order (Mapper.java:240-377). len1/len2 = read.seq.length per record.
// Returns 0, or -1 where the Java would have thrown ArrayIndexOutOfBounds.
int vo_identity_counting(const PairRec* recs, int64_t N, const uint8_t* seq1pool, const int32_t* len1, const int32_t* len2,
                         const uint8_t* genome, int G, int windowSize, int threadNum, float* identityOut, int32_t* coverageOut) {
  struct Item { const MateRec* m; int seqLen; };
  std::vector<Item> list;
  for (int64_t i = 0; i < N; i++)
    for (int k = 0; k < 2; k++)
      if (recs[i].m[k].present) list.push_back({&recs[i].m[k], k == 0 ? len1[i] : len2[i]});
  for (int i = 0; i < G; i++) identityOut[i] = 0.0f, coverageOut[i] = 0;
  if (threadNum < 1) return -1;
  const int64_t size = (int64_t)list.size() / threadNum;
  std::vector<float> sumIdentity((size_t)std::max(G, 1));
  std::vector<int32_t> coverage((size_t)std::max(G, 1));
  for (int t = 0; t < threadNum; t++) {
    const int64_t from = t * size, to = (t == threadNum - 1) ? (int64_t)list.size() : from + size;
    std::fill(sumIdentity.begin(), sumIdentity.end(), 0.0f);
    std::fill(coverage.begin(), coverage.end(), 0);
    for (int64_t n = from; n < to; n++) {
      const MateRec& a = *list[n].m;
      const int seqLen = list[n].seqLen;
      if (seqLen < windowSize) continue;
      const uint8_t* seq1 = seq1pool + a.seq1_offset;
      const int alnLen = a.length;
      int identity = 0;
      const int startNoClip = a.start + a.start2;
      const int start = std::max(0, startNoClip - a.start1);
      const int endNoClip = a.start + a.end2;
      const int end = std::min(G, endNoClip + seqLen - a.end1);
      int alnEndPos = 0, len, alnStartPos = 0;
#define IDC_SEQ(p) ((p) < 0 || (p) >= alnLen ? (oob = true, (uint8_t)0) : seq1[p])
#define IDC_GEN(p) ((p) < 0 || (p) >= G ? (oob = true, (uint8_t)0) : genome[p])
      bool oob = false;
      if (a.start1 >= windowSize) {
        len = windowSize;
      } else {
        len = a.start1;
        if (alnLen > windowSize - a.start1) {
          for (int i = startNoClip; i < startNoClip + windowSize - a.start1 && alnEndPos < alnLen;) {
            const uint8_t c = seq1[alnEndPos];
            if (c != GAP && c == IDC_GEN(i)) {
              identity++, alnEndPos++, len++, i++;
            } else if (c < 'a') {
              alnEndPos++, len++, i++;
            } else {
              do {
                alnEndPos++, len++;
              } while (IDC_SEQ(alnEndPos) >= 'a');
            }
            if (oob) return -1;
          }
          while (alnEndPos < alnLen && seq1[alnEndPos] >= 'a') alnEndPos++, len++;
        } else {
          identity += a.identity;
          len += windowSize - a.start1;
          alnEndPos = alnLen;
        }
      }
      if (start < 0 || start >= G) return -1;
      sumIdentity[start] += (float)identity / (float)len;
      coverage[start]++;
      for (int i = start + 1; i <= end - windowSize; i++) {
        if (i > startNoClip && i <= endNoClip) {
          const uint8_t c = IDC_SEQ(alnStartPos);
          if (c != GAP && c == IDC_GEN(i - 1)) identity--;
          alnStartPos++;
          if (alnStartPos < alnLen)
            while (IDC_SEQ(alnStartPos) >= 'a') alnStartPos++, len--;
        }
        if (i > startNoClip - windowSize && i <= endNoClip - windowSize) {
          const uint8_t c = IDC_SEQ(alnEndPos);
          if (c != GAP && c == IDC_GEN(i + windowSize - 1)) identity++;
          alnEndPos++;
          if (alnEndPos < alnLen)
            while (IDC_SEQ(alnEndPos) >= 'a') alnEndPos++, len++;
        }
        if (oob) return -1;
        sumIdentity[i] += (float)identity / (float)len;
        coverage[i]++;
      }
#undef IDC_SEQ
#undef IDC_GEN
    }
    for (int i = 0; i < G; i++) identityOut[i] += sumIdentity[i], coverageOut[i] += coverage[i];
  }
  return 0;
}

// ConsensusBuildingSimple.run (:98-120)
void vo_consensus(const int32_t* counts, int G, int saveUncovered, int coverageThreshold, const uint8_t* genome,
                  uint8_t* consensus) {
  for (int i = 0; i < G; i++) {
    int maxCount = counts[(size_t)i * 5], maxCountIndex = 0;
    for (int j = 1; j < 4; j++)
      if (counts[(size_t)i * 5 + j] > maxCount) maxCount = counts[(size_t)i * 5 + j], maxCountIndex = j;
    if (maxCount <= coverageThreshold)
      consensus[i] = saveUncovered ? genome[i] : GAP;
    else
      consensus[i] = T.iToN[maxCountIndex];
  }
}

// BamPrinter.setCigar (BamPrinter.java:121-171): CIGAR ops as (len<<4|op) with op M=0,I=1,D=2,S=4; returns nOps,
// *xn = mismatches (XN tag). Note the reference emits a leading zero-length M when the alignment starts with I/D.
int vo_cigar(const uint8_t* seq1, int length, int start1, int end1, int readLen, int identity, uint32_t* ops, int cap,
             int* xn) {
  int n = 0;
  auto add = [&](int len, int op) {
    if (n < cap) ops[n] = ((uint32_t)len << 4) | (uint32_t)op;
    n++;
  };
  if (start1 > 0) add(start1, 4);
  int op = 0, len = 0, mNum = 0;
  for (int i = 0; i < length; i++) {
    uint8_t ch = seq1[i];
    if (ch == GAP) {
      if (op == 2) len++;
      else add(len, op), op = 2, len = 1;
    } else if (ch < 'a') {
      if (op == 0) len++;
      else add(len, op), op = 0, len = 1;
      mNum++;
    } else {
      if (op == 1) len++;
      else add(len, op), op = 1, len = 1;
    }
  }
  add(len, op);
  if (end1 < readLen) add(readLen - end1, 4);
  *xn = mNum - identity;
  return n;
}

void vo_revcomp(const uint8_t* seq, int L, uint8_t* out) {
  std::string r = getComplement(std::string((const char*)seq, L));
  memcpy(out, r.data(), L);
}

// RandomModel.genModel + computeCuttofs with a seeded RNG. genomes = contigs of the reference when
// fragmented, else the whole sequence (RandomModel.java:88-98). cutoffs has maxReadLen+1 entries.
// The random model's own counter uses lowCoef = 1/highCoef (KMerCounter.java:51-55, Q3).
// countsOut (optional): the sorted scores, [size][readNum]; readsOut (optional): the readNum reads of the first length
// generated (maxReadLen), readNum x maxReadLen bytes.
void vo_random_model(const uint8_t* seq, int G, const int32_t* contigEnds, int nC, int isFragmented, int K, float coef,
                     int order, int readNum, int minReadLen, int maxReadLen, int step, float pValue, int indexThreadNum,
                     uint64_t seed, int msa, const uint8_t* msaRows, int nRefs, int alnLength, int32_t* cutoffs,
                     int32_t* countsOut, uint8_t* readsOut) {
  std::string s((const char*)seq, G);
  KMerCounter kc;
  kc.K = K;
  kc.highCoef = coef;
  kc.lowCoef = 1.0f / coef;
  Index index;
  std::vector<std::string> genomes;
  int from = 0, to = G;
  if (msa) {
    // RandomModelMSA: model over all ungapped references, scoring against the MSA index on [0, alnLength)
    std::vector<std::string> alns(nRefs);
    for (int r = 0; r < nRefs; r++) {
      alns[r].assign((const char*)msaRows + (size_t)r * alnLength, alnLength);
      std::string u;
      for (char ch : alns[r])
        if (ch != '-') u.push_back(ch);
      genomes.push_back(u);
    }
    index = buildAlnIndex(alns, K);
    to = alnLength;
  } else {
    index = buildIndexPara(s, K, indexThreadNum);
    if (isFragmented) {
      for (int i = 0, j = 0; i < nC; i++) {
        genomes.push_back(s.substr(j, contigEnds[i] - j));
        j = contigEnds[i];
      }
    } else {
      genomes.push_back(s);
    }
  }
  MarkovModel model(genomes, order);
  int size = (maxReadLen - minReadLen) / step + 1;
  std::vector<std::vector<int>> counts(size, std::vector<int>(readNum));
  Rng rnd(seed);
  for (int len = maxReadLen, i = size - 1; i >= 0; len -= step, i--) {
    for (int r = 0; r < readNum; r++) {
      const std::string read = model.generate(len, rnd);
      if (readsOut && len == maxReadLen) memcpy(readsOut + (size_t)r * maxReadLen, read.data(), maxReadLen);
      counts[i][r] = kc.computeKMerCount(read, index, from, to);
    }
    std::sort(counts[i].begin(), counts[i].end());
    if (countsOut) memcpy(countsOut + (size_t)i * readNum, counts[i].data(), (size_t)readNum * 4);
  }
  // computeCuttofs (KMerCounterBase.java:229-251); note counts here is [size][readNum] so num = size (sic: the
  // Java passes model.counts = int[size][readNum] and uses counts.length as `num`).
  int num = size;
  std::vector<int> interp(size);
  for (int i = 0; i < size; i++) {
    int idx = (int)std::floor(num * (1.0f - pValue) + 0.5f);
    interp[i] = counts[i][idx];
  }
  for (int i = 0; i <= minReadLen; i++) cutoffs[i] = interp[0];
  for (int i = maxReadLen; i > minReadLen; i--) {
    int right = size - (maxReadLen - i) / step - 1;
    if (right == 0) {
      cutoffs[i] = interp[0];
    } else {
      int distToRight = (maxReadLen - i) % step;
      int distToLeft = step - distToRight;
      cutoffs[i] = (distToRight * interp[right - 1] + distToLeft * interp[right]) / step;
    }
  }
}

void vo_cutoffs(const uint8_t* seq, int G, const int32_t* contigEnds, int nC, int isFragmented, int K, float coef,
                int order, int readNum, int minReadLen, int maxReadLen, int step, float pValue, int indexThreadNum,
                uint64_t seed, int msa, const uint8_t* msaRows, int nRefs, int alnLength, int32_t* cutoffs) {
  vo_random_model(seq, G, contigEnds, nC, isFragmented, K, coef, order, readNum, minReadLen, maxReadLen, step, pValue,
                  indexThreadNum, seed, msa, msaRows, nRefs, alnLength, cutoffs, nullptr, nullptr);
}

}  // extern "C"
