/*
 * pose_driver.cc -- drives the ORBmatcher overloads that PROJECT map points themselves (general camera poses, Sim3
 * transforms) and the two Fuse overloads + SearchBySim3, and prints everything they write.  One source, two builds:
 *
 *   product   g++ -I viorb_b200/host ... -lviorb_b200           ORB_SLAM2::ORBmatcher = the CUDA drop-in (needs a GPU)
 *   reference oracle/refbuild/Makefile -> oracle/_ref/pose_driver_ref
 *                                                                ORB_SLAM2::ORBmatcher = /root/reference/src/ORBmatcher.cc,
 *                                                                compiled unmodified (namespace renamed on the command line)
 *
 * Both builds use the same stand-in Frame / KeyFrame / MapPoint classes (viorb_b200/host/orbslam_compat.h) and the same
 * deterministic scenario, so the two outputs must be identical line by line (tests/test_cpp_shims.py; the reference
 * build's output is also committed as tests/golden/ref_pose_driver.txt).  This is what pins the host arithmetic of the
 * shims (R*x+t, -R.t()*t, M/s, dot, norm: OpenCV's small-matrix conventions) on the reference's own code.
 *
 * usage: pose_driver <keys.bin> <out.txt>     keys.bin = int32 n1, n1 x KeyPoint(28 B), n1 x 32 B, int32 n2, ... (tests write it)
 */
#include "pose_scenarios.h"

static std::map<MapPoint*, int> g_ids;
static int id_of(MapPoint* p) {
    if (!p) return -1;
    std::map<MapPoint*, int>::iterator it = g_ids.find(p);
    return it == g_ids.end() ? -2 : it->second;
}
static void name_points(std::vector<MapPoint>& store, int base) {
    for (size_t i = 0; i < store.size(); i++) g_ids[&store[i]] = base + (int)i;
}
static void dump_slots(FILE* out, const char* tag, const std::vector<MapPoint*>& v) {
    fprintf(out, "%s", tag);
    for (size_t i = 0; i < v.size(); i++) fprintf(out, " %d", id_of(v[i]));
    fprintf(out, "\n");
}
static void dump_points(FILE* out, const char* tag, std::vector<MapPoint>& store) {
    fprintf(out, "%s", tag);
    for (size_t i = 0; i < store.size(); i++)
        fprintf(out, " %d:%d:%d:%d", (int)store[i].isBad(), id_of(store[i].mpReplaced), store[i].Observations(), (int)store[i].mObservations.size());
    fprintf(out, "\n");
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: pose_driver keys.bin out.txt\n"); return 2; }
    FILE* fin = fopen(argv[1], "rb");
    Keys A, B;
    if (!fin || !read_keys(fin, A) || !read_keys(fin, B)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(fin);
    FILE* out = fopen(argv[2], "w");
    if (!out) return 2;
    const int NQ = 700;

    for (int round = 0; round < 3; round++) {
        const Pose cur = random_pose(0.12, 0.8);
        /* ---- SearchByProjection(Frame& Cur, const Frame& Last, th, bMono), src/ORBmatcher.cc:1328-1471 ---- */
        for (int variant = 0; variant < 3; variant++) {
            Frame Cur;
            setup_frame(Cur, A, cur);
            std::vector<MapPoint> store(NQ), held(40);
            g_ids.clear(); name_points(store, 0); name_points(held, 100000);
            for (int i = 0; i < 40; i++) { held[i].nObs = (int)(rnd() % 3); Cur.mvpMapPoints[(rnd() % Cur.N)] = &held[i]; }
            Frame Last;
            Pose last = cur;
            last.t[2] += variant == 1 ? 1.5 : (variant == 2 ? -1.5 : 0.05);          /* forward / backward / neither (:1346-1349) */
            Last.N = NQ; Last.mTcw = mat44(last);
            Last.mvpMapPoints.assign(NQ, static_cast<MapPoint*>(NULL));
            Last.mvbOutlier.assign(NQ, false);
            Last.mvKeys.resize(NQ); Last.mvKeysUn.resize(NQ);
            for (int i = 0; i < NQ; i++) {
                const int k = i < 60 ? (int)(rnd() % 30) : (int)(rnd() % Cur.N);
                make_point(store[i], A, k, cur, 5.0f, 0.3f);
                if (rnd() % 50 == 0) store[i].worldPos.at<float>(2) = -store[i].worldPos.at<float>(2);      /* behind the camera */
                Last.mvpMapPoints[i] = (rnd() % 12) == 0 ? NULL : &store[i];
                Last.mvbOutlier[i] = (rnd() % 15) == 0;
                Last.mvKeys[i] = A.k[k]; Last.mvKeysUn[i] = A.k[k];
                Last.mvKeysUn[i].angle = A.k[k].angle + ((rnd() % 6) == 0 ? 150.f : 3.f * rndf());
            }
            ORBmatcher matcher(0.9f, true);
            const int n = matcher.SearchByProjection(Cur, Last, variant == 0 ? 15.f : 7.f, round == 2 && variant == 0);
            fprintf(out, "last_frame round %d variant %d n %d\n", round, variant, n);
            dump_slots(out, "last_frame slots", Cur.mvpMapPoints);
        }
        /* ---- SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist), :1473-1600 ---- */
        {
            Frame Cur;
            setup_frame(Cur, A, cur);
            std::vector<MapPoint> store(NQ), held(30);
            g_ids.clear(); name_points(store, 0); name_points(held, 100000);
            for (int i = 0; i < 30; i++) Cur.mvpMapPoints[(rnd() % Cur.N)] = &held[i];
            KeyFrame kf;
            kf.N = NQ; kf.mvKeysUn.resize(NQ); kf.mapPoints.assign(NQ, static_cast<MapPoint*>(NULL));
            std::set<MapPoint*> found;
            for (int i = 0; i < NQ; i++) {
                const int k = i < 60 ? (int)(rnd() % 30) : (int)(rnd() % Cur.N);
                make_point(store[i], A, k, cur, 8.0f, 0.3f);
                kf.mapPoints[i] = (rnd() % 11) == 0 ? NULL : &store[i];
                if ((rnd() % 13) == 0) found.insert(&store[i]);
                kf.mvKeysUn[i] = A.k[k];
                kf.mvKeysUn[i].angle = A.k[k].angle + ((rnd() % 5) == 0 ? 120.f : 2.f * rndf());
            }
            ORBmatcher matcher(0.9f, true);
            const int n = matcher.SearchByProjection(Cur, &kf, found, round == 1 ? 3.f : 10.f, round == 1 ? 64 : 100);
            fprintf(out, "reloc round %d n %d\n", round, n);
            dump_slots(out, "reloc slots", Cur.mvpMapPoints);
        }
        /* ---- SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), :290-403 ---- */
        {
            KeyFrame kf;
            setup_keyframe(kf, A, cur);
            std::vector<MapPoint> store(NQ);
            g_ids.clear(); name_points(store, 0);
            std::vector<MapPoint*> vpPoints(NQ), vpMatched(kf.N, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < NQ; i++) {
                const int k = i < 60 ? (int)(rnd() % 30) : (int)(rnd() % kf.N);
                make_point(store[i], A, k, cur, 8.0f, 1.3f);
                vpPoints[i] = &store[i];
            }
            for (int i = 0; i < 25; i++) vpMatched[rnd() % kf.N] = &store[rnd() % NQ];
            const double s = 0.8 + 0.5 * rndf();
            ORBmatcher matcher(0.75f, true);
            const int n = matcher.SearchByProjection(&kf, mat44(cur, s), vpPoints, vpMatched, 10);
            fprintf(out, "sim3_projection round %d n %d\n", round, n);
            dump_slots(out, "sim3_projection matched", vpMatched);
        }
        /* ---- Fuse(KeyFrame*, const vector<MapPoint*>&, th), :825-976 ---- */
        {
            KeyFrame kf;
            setup_keyframe(kf, A, cur);
            std::vector<MapPoint> store(NQ), inKF(kf.N);
            g_ids.clear(); name_points(store, 0); name_points(inKF, 100000);
            for (int k = 0; k < kf.N; k += 2) {                     /* half of the key frame's keypoints already hold a point */
                inKF[k].nObs = 0;
                inKF[k].AddObservation(&kf, k);
                for (int e = 0; e < (int)(rnd() % 4); e++) inKF[k].nObs++;
                inKF[k].bad = (rnd() % 31) == 0;
                kf.mapPoints[k] = &inKF[k];
            }
            std::vector<MapPoint*> vp(NQ);
            for (int i = 0; i < NQ; i++) {
                const int k = i < 80 ? (int)(rnd() % 40) : (int)(rnd() % kf.N);
                make_point(store[i], A, k, cur, 3.0f, 1.3f);
                store[i].nObs = (int)(rnd() % 4);
                vp[i] = (rnd() % 29) == 0 ? NULL : &store[i];
            }
            for (int i = 0; i < 20; i++) {                          /* some candidates are already in the key frame */
                const int k = (int)(rnd() % kf.N);
                if (!kf.mapPoints[k]) { kf.mapPoints[k] = &store[i]; store[i].AddObservation(&kf, k); }
            }
            ORBmatcher matcher(0.6f, true);
            const int n = matcher.Fuse(&kf, vp, 3.0f);
            fprintf(out, "fuse round %d n %d\n", round, n);
            dump_slots(out, "fuse keyframe", kf.mapPoints);
            dump_points(out, "fuse candidates", store);
            dump_points(out, "fuse residents", inKF);
        }
        /* ---- Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint), :978-1100 ---- */
        {
            KeyFrame kf;
            setup_keyframe(kf, A, cur);
            std::vector<MapPoint> store(NQ), inKF(kf.N);
            g_ids.clear(); name_points(store, 0); name_points(inKF, 100000);
            for (int k = 0; k < kf.N; k += 3) { inKF[k].bad = (rnd() % 19) == 0; kf.mapPoints[k] = &inKF[k]; }
            std::vector<MapPoint*> vp(NQ), repl(NQ, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < NQ; i++) {
                const int k = i < 80 ? (int)(rnd() % 40) : (int)(rnd() % kf.N);
                make_point(store[i], A, k, cur, 3.0f, 1.3f);
                vp[i] = &store[i];
            }
            const double s = 0.8 + 0.5 * rndf();
            ORBmatcher matcher(0.8f, true);
            const int n = matcher.Fuse(&kf, mat44(cur, s), vp, 4.0f, repl);
            fprintf(out, "fuse_sim3 round %d n %d\n", round, n);
            dump_slots(out, "fuse_sim3 keyframe", kf.mapPoints);
            dump_slots(out, "fuse_sim3 replace", repl);
            dump_points(out, "fuse_sim3 candidates", store);
        }
        /* ---- SearchBySim3(KeyFrame*, KeyFrame*, vpMatches12, s12, R12, t12, th), :1102-1326 ---- */
        {
            const Pose p1 = cur;
            Pose p2 = cur;
            p2.t[0] -= 0.5372;                                       /* the right camera of the stereo pair */
            KeyFrame k1, k2;
            setup_keyframe(k1, A, p1);
            setup_keyframe(k2, B, p2);
            std::vector<MapPoint> s1(k1.N), s2(k2.N);
            g_ids.clear(); name_points(s1, 0); name_points(s2, 100000);
            for (int i = 0; i < k1.N; i++) {
                make_point(s1[i], A, i, p1, 0.5f, 0.2f);
                if (rnd() % 10 < 8) { k1.mapPoints[i] = &s1[i]; s1[i].AddObservation(&k1, i); }
            }
            for (int i = 0; i < k2.N; i++) {
                make_point(s2[i], B, i, p2, 0.5f, 0.2f);
                if (rnd() % 10 < 8) { k2.mapPoints[i] = &s2[i]; s2[i].AddObservation(&k2, i); }
            }
            std::vector<MapPoint*> m12(k1.N, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < 15; i++) { const int a = rnd() % k1.N, b = rnd() % k2.N; if (k2.mapPoints[b]) m12[a] = k2.mapPoints[b]; }
            /* T12 = T1w * Tw2 with a scale a little off 1 */
            double R12[9], t12[3];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) R12[3 * r + c] = p1.R[3 * r] * p2.R[3 * c] + p1.R[3 * r + 1] * p2.R[3 * c + 1] + p1.R[3 * r + 2] * p2.R[3 * c + 2];
            for (int r = 0; r < 3; r++) t12[r] = p1.t[r] - (R12[3 * r] * p2.t[0] + R12[3 * r + 1] * p2.t[1] + R12[3 * r + 2] * p2.t[2]);
            const float s12 = 0.97f + 0.06f * rndf();
            ORBmatcher matcher(0.75f, true);
            const int n = matcher.SearchBySim3(&k1, &k2, m12, s12, mat33(R12), mat31(t12), 7.5f);
            fprintf(out, "search_by_sim3 round %d n %d\n", round, n);
            dump_slots(out, "search_by_sim3 matches12", m12);
        }
        /* ---- SearchForTriangulation(KeyFrame*, KeyFrame*, F12, pairs, bOnlyStereo), :657-823: the epipole comes from poses ---- */
        {
            const Pose p1 = cur;
            Pose p2 = cur;
            p2.t[0] -= 0.5372; p2.t[2] += 0.3 * (round - 1);
            KeyFrame k1, k2;
            setup_keyframe(k1, A, p1);
            setup_keyframe(k2, B, p2);
            std::vector<MapPoint> have(64);
            for (int i = 0; i < 64; i++) { k1.mapPoints[rnd() % k1.N] = &have[i]; k2.mapPoints[rnd() % k2.N] = &have[i]; }
            for (int i = 0; i < k1.N; i++) k1.mFeatVec[(unsigned)(A.k[i].pt.y / 24) * 3 + 5].push_back(i);
            for (int i = 0; i < k2.N; i++) k2.mFeatVec[(unsigned)(B.k[i].pt.y / 24) * 3 + 5].push_back(i);
            cv::Mat F12 = cv::Mat(3, 3, CV_32F);                     /* rectified pair: l = x1' F12 = [0, 1, -y1] */
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) F12.at<float>(r, c) = 0.f;
            F12.at<float>(1, 2) = -1; F12.at<float>(2, 1) = 1;
            std::vector<std::pair<size_t, size_t> > pairs;
            ORBmatcher matcher(0.6f, round != 0);
            const int n = matcher.SearchForTriangulation(&k1, &k2, F12, pairs, round == 2);
            fprintf(out, "triangulation round %d n %d\n", round, n);
            fprintf(out, "triangulation pairs");
            for (size_t i = 0; i < pairs.size(); i++) fprintf(out, " %zu:%zu", pairs[i].first, pairs[i].second);
            fprintf(out, "\n");
        }
    }
    fclose(out);
    return 0;
}
