// ref_harness.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own, UNMODIFIED line_descriptor
// sources (/root/reference/src/line_descriptor/src/{LSDDetector_custom,binary_descriptor_custom,binary_descriptor_matcher}.cpp),
// compiled against oracle/ref_shim/cvshim.hpp by oracle/build_ref.py into oracle/_ref/libref_line_descriptor.so.
// Used only to generate golden vectors (tests/golden/make_golden_lbd.py) and by the CPU tests that pin the oracle's
// restatement of KeyLine fill / computeLBD / knnMatch to the reference's compiled code.
#include "precomp_custom.hpp"

using namespace cv;
using namespace cv::line_descriptor;

extern "C" {

// LSDDetectorC::detect (LSDDetector_custom.cpp:105-215) on `lines` (injected in place of cv::LineSegmentDetector's
// output) followed by BinaryDescriptor::compute (binary_descriptor_custom.cpp:524-687).
// keyl: [n][10] = startX, startY, endX, endY, lineLength, numOfPixels, angle, response, size, class_id
__attribute__((visibility("default"))) int ref_keylines_lbd(const uint8_t *gray, int H, int W, const float *lines, int n, float *keyl,
                                                            float *desc72, uint8_t *desc32)
{
    try {
        Mat img(H, W, CV_8UC1, (void *)gray);
        std::vector<Vec4f> &inj = LineSegmentDetector::injected();
        inj.clear();
        for (int i = 0; i < n; ++i) inj.push_back(Vec4f(lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]));
        std::vector<KeyLine> kls;
        Ptr<LSDDetectorC> det = LSDDetectorC::createLSDDetectorC();
        det->detect(img, kls, 2, 1);
        if ((int)kls.size() != n) return -1;
        for (int i = 0; i < n; ++i) {
            const KeyLine &k = kls[i];
            float *o = keyl + 10 * i;
            o[0] = k.startPointX; o[1] = k.startPointY; o[2] = k.endPointX; o[3] = k.endPointY; o[4] = k.lineLength;
            o[5] = (float)k.numOfPixels; o[6] = k.angle; o[7] = k.response; o[8] = k.size; o[9] = (float)k.class_id;
        }
        if (n == 0) return 0;
        Ptr<BinaryDescriptor> bd = BinaryDescriptor::createBinaryDescriptor();
        Mat d32, d72;
        bd->compute(img, kls, d32, false);
        bd->compute(img, kls, d72, true);
        if (d32.rows != n || d32.cols != 32 || d72.rows != n || d72.cols != 72) return -2;
        for (int i = 0; i < n; ++i) {
            memcpy(desc32 + 32 * i, d32.ptr<uchar>(i), 32);
            memcpy(desc72 + 72 * i, d72.ptr<float>(i), 72 * sizeof(float));
        }
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_keylines_lbd: %s\n", e.what());
        return -3;
    }
}

// BinaryDescriptorMatcher::knnMatch(query, train, matches, k) (binary_descriptor_matcher.cpp:258-335).
// idx/dist [nq][k], -1 where the matcher returned fewer than k neighbours.
__attribute__((visibility("default"))) int ref_knn_match(const uint8_t *q, int nq, const uint8_t *m, int nm, int k, int *idx, int *dist)
{
    try {
        Mat Q(nq, 32, CV_8UC1, (void *)q), M(nm, 32, CV_8UC1, (void *)m);
        Ptr<BinaryDescriptorMatcher> bm = BinaryDescriptorMatcher::createBinaryDescriptorMatcher();
        std::vector<std::vector<DMatch> > matches;
        bm->knnMatch(Q, M, matches, k);
        for (int i = 0; i < nq * k; ++i) idx[i] = dist[i] = -1;
        for (size_t r = 0; r < matches.size(); ++r)
            for (size_t j = 0; j < matches[r].size() && (int)j < k; ++j) {
                const DMatch &d = matches[r][j];
                if (d.queryIdx < 0 || d.queryIdx >= nq) continue;
                idx[d.queryIdx * k + j] = d.trainIdx;
                dist[d.queryIdx * k + j] = (int)d.distance;
            }
        return (int)matches.size();
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_knn_match: %s\n", e.what());
        return -3;
    }
}

}  // extern "C"
