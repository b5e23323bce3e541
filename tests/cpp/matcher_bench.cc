/*
 * matcher_bench.cc -- per-call wall time of the call surfaces north_star names, through the C++ classes exactly as Frame /
 * Tracking / LocalMapping call them: ORBextractor::operator(), Frame::ComputeStereoMatches, the four
 * ORBmatcher::SearchByProjection overloads, SearchForTriangulation, SearchByBoW, Fuse, DescriptorDistance.
 * One source, two builds (like pose_driver.cc):
 *
 *   product    -I viorb_b200/host -lviorb_b200                 the CUDA drop-in (host buffers in, host results out: every call
 *                                                               includes its H2D/D2H copies and synchronisation)
 *   reference  oracle/refbuild/Makefile -> oracle/_ref/matcher_bench_ref
 *                                                               /root/reference/src/ORBextractor.cc + ORBmatcher.cc + Frame.cc
 *                                                               members compiled unmodified, -O3, one thread (how VIORB runs them)
 *
 * KITTI-shape scenario (BASELINE configs[1]): 1241x376 stereo pair, 2000 features per image, ~700-2000 map points per
 * search.  Prints one JSON object per surface: {"surface": ..., "us_median": ..., "us_min": ..., "n": result count}.
 * usage: matcher_bench <pair.bin> <reps>        pair.bin = left image then right image, 376 x 1241 bytes each
 */
#include <algorithm>
#include <chrono>
#include <functional>
#include <string>

#include "ORBextractor.h"
#include "pose_scenarios.h"

static double now_us() {
    using namespace std::chrono;
    return duration<double, std::micro>(steady_clock::now().time_since_epoch()).count();
}

/* setup() rebuilds the state the call mutates (untimed), call() is timed */
static void bench(const char* name, int reps, const std::function<void()>& setup, const std::function<int()>& call) {
    std::vector<double> t;
    int n = 0;
    for (int r = 0; r < reps + 2; r++) {
        setup();
        const double t0 = now_us();
        n = call();
        const double t1 = now_us();
        if (r >= 2) t.push_back(t1 - t0);           /* two warm-up calls */
    }
    std::sort(t.begin(), t.end());
    printf("{\"surface\": \"%s\", \"us_median\": %.1f, \"us_min\": %.1f, \"reps\": %d, \"n\": %d}\n", name, t[t.size() / 2], t[0], reps, n);
    fflush(stdout);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: matcher_bench pair.bin reps\n"); return 2; }
    const int reps = atoi(argv[2]);
    cv::Mat left(H, W, CV_8UC1), right(H, W, CV_8UC1);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(left.data, 1, (size_t)H * W, f) != (size_t)H * W || fread(right.data, 1, (size_t)H * W, f) != (size_t)H * W) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    fclose(f);

    ORBextractor exL(2000, 1.2f, NLEVELS, 20, 7), exR(2000, 1.2f, NLEVELS, 20, 7);
    Keys A, B;
    /* ---- ORBextractor::operator() (src/ORBextractor.cc:1043-1105) ---- */
    bench("ORBextractor::operator() 1241x376 nf=2000", reps, [] {}, [&] { exL(left, cv::Mat(), A.k, A.d); return (int)A.k.size(); });
    bench("ORBextractor::operator() + mvImagePyramid read", reps, [] {}, [&] {
        exL(left, cv::Mat(), A.k, A.d);
        int rows = 0;
        for (int l = 0; l < NLEVELS; l++) rows += exL.mvImagePyramid[l].rows;
        return rows;
    });
    exL(left, cv::Mat(), A.k, A.d);
    exR(right, cv::Mat(), B.k, B.d);

    /* ---- Frame::ComputeStereoMatches (src/Frame.cc:646-820) ---- */
    {
        Frame F;
        F.N = (int)A.k.size();
        F.mvKeys = A.k; F.mvKeysUn = A.k; F.mvKeysRight = B.k; F.mDescriptors = A.d; F.mDescriptorsRight = B.d;
        F.mvScaleFactors = scale_factors();
        F.mvInvScaleFactors.resize(NLEVELS);
        for (int l = 0; l < NLEVELS; l++) F.mvInvScaleFactors[l] = 1.0f / F.mvScaleFactors[l];
        F.mbf = BF; F.mb = BF / FX;
        F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
        bench("Frame::ComputeStereoMatches", reps, [] {}, [&] {
            F.ComputeStereoMatches();
            int n = 0;
            for (int i = 0; i < F.N; i++) n += F.mvuRight[i] >= 0;
            return n;
        });
    }

    const Pose cur = random_pose(0.12, 0.8);
    const int NQ = 2000;
    std::vector<MapPoint> store(NQ);
    std::vector<int> pick(NQ);
    for (int i = 0; i < NQ; i++) {
        pick[i] = i < 100 ? (int)(rnd() % 50) : (int)(rnd() % A.k.size());
        make_point(store[i], A, pick[i], cur, 5.0f, 0.3f);
        store[i].bad = false;
    }

    /* ---- SearchByProjection(Frame&, const vector<MapPoint*>&, th) (:45-129): Tracking::SearchLocalPoints ---- */
    {
        Frame F;
        setup_frame(F, A, cur);
        std::vector<MapPoint*> vp(NQ);
        for (int i = 0; i < NQ; i++) {
            MapPoint& p = store[i];
            p.mbTrackInView = true;
            p.mTrackProjX = A.k[pick[i]].pt.x + (rndf() - 0.5f) * 4; p.mTrackProjY = A.k[pick[i]].pt.y + (rndf() - 0.5f) * 4;
            p.mTrackProjXR = p.mTrackProjX - 20;
            p.mnTrackScaleLevel = std::min(7, std::max(0, A.k[pick[i]].octave + (int)(rnd() % 3) - 1));
            p.mTrackViewCos = 0.99f + 0.01f * rndf();
            vp[i] = &p;
        }
        ORBmatcher matcher(0.8f, true);
        bench("SearchByProjection(Frame, vector<MapPoint*>, th) 2000 points", reps,
              [&] { F.mvpMapPoints.assign(F.N, static_cast<MapPoint*>(NULL)); }, [&] { return matcher.SearchByProjection(F, vp, 3.0f); });
    }
    /* ---- SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (:1328-1471): Tracking::TrackWithMotionModel ---- */
    {
        Frame Cur, Last;
        setup_frame(Cur, A, cur);
        Pose last = cur;
        last.t[2] += 0.05;
        Last.N = NQ; Last.mTcw = mat44(last);
        Last.mvpMapPoints.resize(NQ); Last.mvbOutlier.assign(NQ, false);
        Last.mvKeys.resize(NQ); Last.mvKeysUn.resize(NQ);
        for (int i = 0; i < NQ; i++) { Last.mvpMapPoints[i] = &store[i]; Last.mvKeys[i] = A.k[pick[i]]; Last.mvKeysUn[i] = A.k[pick[i]]; }
        ORBmatcher matcher(0.9f, true);
        bench("SearchByProjection(Frame, Frame, th, bMono) 2000 points", reps,
              [&] { Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL)); }, [&] { return matcher.SearchByProjection(Cur, Last, 15.f, false); });
    }
    /* ---- SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist) (:1473-1600): Tracking::Relocalization ---- */
    {
        Frame Cur;
        setup_frame(Cur, A, cur);
        KeyFrame kf;
        kf.N = NQ; kf.mvKeysUn.resize(NQ); kf.mapPoints.resize(NQ);
        for (int i = 0; i < NQ; i++) { kf.mapPoints[i] = &store[i]; kf.mvKeysUn[i] = A.k[pick[i]]; }
        std::set<MapPoint*> found;
        ORBmatcher matcher(0.9f, true);
        bench("SearchByProjection(Frame, KeyFrame*, set, th, ORBdist) 2000 points", reps,
              [&] { Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL)); }, [&] { return matcher.SearchByProjection(Cur, &kf, found, 10.f, 100); });
    }
    /* ---- SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (:290-403): LoopClosing::ComputeSim3 ---- */
    {
        KeyFrame kf;
        setup_keyframe(kf, A, cur);
        std::vector<MapPoint*> vpPoints(NQ), vpMatched;
        for (int i = 0; i < NQ; i++) vpPoints[i] = &store[i];
        ORBmatcher matcher(0.75f, true);
        bench("SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) 2000 points", reps,
              [&] { vpMatched.assign(kf.N, static_cast<MapPoint*>(NULL)); }, [&] { return matcher.SearchByProjection(&kf, mat44(cur, 1.05), vpPoints, vpMatched, 10); });
    }
    /* ---- SearchForTriangulation (:657-823): LocalMapping::CreateNewMapPoints, and SearchByBoW (:159-288) ---- */
    {
        Pose p2 = cur;
        p2.t[0] -= 0.5372;
        KeyFrame k1, k2;
        setup_keyframe(k1, A, cur);
        setup_keyframe(k2, B, p2);
        for (int i = 0; i < k1.N; i++) k1.mFeatVec[(unsigned)(A.k[i].pt.y / 24) * 3 + 5].push_back(i);
        for (int i = 0; i < k2.N; i++) k2.mFeatVec[(unsigned)(B.k[i].pt.y / 24) * 3 + 5].push_back(i);
        cv::Mat F12(3, 3, CV_32F);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) F12.at<float>(r, c) = 0.f;
        F12.at<float>(1, 2) = -1; F12.at<float>(2, 1) = 1;
        std::vector<std::pair<size_t, size_t> > pairs;
        ORBmatcher matcher(0.6f, false);
        bench("SearchForTriangulation(KeyFrame*, KeyFrame*, F12, pairs, false) 2x2000 keypoints", reps, [] {},
              [&] { return matcher.SearchForTriangulation(&k1, &k2, F12, pairs, false); });

        std::vector<MapPoint> pool(k1.N);
        for (int i = 0; i < k1.N; i++) k1.mapPoints[i] = &pool[i];
        Frame F;
        F.N = k2.N; F.mvKeys = B.k; F.mvKeysUn = B.k; F.mDescriptors = B.d; F.mFeatVec = k2.mFeatVec;
        std::vector<MapPoint*> out;
        ORBmatcher bow(0.75f, true);
        bench("SearchByBoW(KeyFrame*, Frame&, matches) 2x2000 keypoints", reps, [] {}, [&] { return bow.SearchByBoW(&k1, F, out); });
    }
    /* ---- Fuse(KeyFrame*, vector<MapPoint*>, th) (:825-976): LocalMapping::SearchInNeighbors ---- */
    {
        KeyFrame kf;
        std::vector<MapPoint> cand, inKF;
        std::vector<MapPoint*> vp;
        ORBmatcher matcher(0.6f, true);
        bench("Fuse(KeyFrame*, vector<MapPoint*>, th) 2000 points", reps,
              [&] {
                  kf = KeyFrame();
                  setup_keyframe(kf, A, cur);
                  cand = store;
                  inKF.assign(kf.N, MapPoint());
                  for (int k = 0; k < kf.N; k += 2) { inKF[k].nObs = 0; inKF[k].AddObservation(&kf, k); kf.mapPoints[k] = &inKF[k]; }
                  vp.resize(NQ);
                  for (int i = 0; i < NQ; i++) vp[i] = &cand[i];
              },
              [&] { return matcher.Fuse(&kf, vp, 3.0f); });
    }
    /* ---- DescriptorDistance (:1648-1664) ---- */
    {
        /* row headers made once: Mat::row() (a refcounted header in either cv) is the caller's cost, not the function's */
        const int nr = std::min(A.d.rows, 2000);
        std::vector<cv::Mat> rows(nr);
        for (int i = 0; i < nr; i++) rows[i] = A.d.row(i);
        bench("DescriptorDistance x 100000", reps, [] {}, [&] {
            int acc = 0;
            for (int i = 0; i < 100000; i++) acc += ORBmatcher::DescriptorDistance(rows[i % nr], rows[(i * 7 + 3) % nr]);
            return acc & 0xffff;
        });
    }
    return 0;
}
