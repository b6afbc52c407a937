"""Build step of oracle/_ref/libft_ref_frame.so (TEST INFRASTRUCTURE ONLY): copies the text of the hot-path functions of the
reference's Frame.cc / MapPoint.cc / ORBmatcher.cc / camera models, verbatim, from where the files lie under the
reference tree into oracle/_ref/gen_frame_fns.inc (git-ignored; nothing of it is stored in the repository). Each range
is checked against the signature expected on its first line, so a different reference revision fails loudly."""
import os
import sys

RANGES = [  # file, first line, last line, text expected on the first line
    ("src/Frame.cc", 409, 440, "void Frame::AssignFeaturesToGrid()"),
    ("src/Frame.cc", 536, 610, "bool Frame::isInFrustum(MapPoint *pMP, float viewingCosLimit)"),
    ("src/Frame.cc", 681, 747, "vector<size_t> Frame::GetFeaturesInArea("),
    ("src/Frame.cc", 749, 759, "bool Frame::PosInGrid("),
    ("src/Frame.cc", 771, 804, "void Frame::UndistortKeyPoints()"),
    ("src/Frame.cc", 806, 833, "void Frame::ComputeImageBounds(const cv::Mat &imLeft)"),
    ("src/Frame.cc", 835, 1005, "void Frame::ComputeStereoMatches()"),
    ("src/Frame.cc", 1065, 1086, "void Frame::ComputeStereoFromRGBD("),
    ("src/Frame.cc", 1231, 1271, "void Frame::ComputeStereoFishEyeMatches()"),
    ("src/Frame.cc", 1308, 1382, "bool Frame::isInFrustumChecks("),
    ("src/MapPoint.cc", 502, 512, "float MapPoint::GetMinDistanceInvariance()"),
    ("src/MapPoint.cc", 531, 546, "int MapPoint::PredictScale(const float &currentDist, Frame* pF)"),
    ("src/ORBmatcher.cc", 49, 312, "int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints"),
    ("src/ORBmatcher.cc", 314, 320, "float ORBmatcher::RadiusByViewingCos("),
    ("src/ORBmatcher.cc", 322, 524, "int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F"),
    ("src/ORBmatcher.cc", 1775, 2085, "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame"),
    ("src/ORBmatcher.cc", 2210, 2251, "void ORBmatcher::ComputeThreeMaxima("),
    ("src/ORBmatcher.cc", 2256, 2272, "int ORBmatcher::DescriptorDistance("),
    ("src/CameraModels/Pinhole.cpp", 43, 49, "Eigen::Vector2f Pinhole::project(const Eigen::Vector3f &v3D)"),
    ("src/CameraModels/KannalaBrandt8.cpp", 67, 93, "Eigen::Vector2f KannalaBrandt8::project(const Eigen::Vector3f &v3D)"),
    ("src/CameraModels/KannalaBrandt8.cpp", 111, 114, "Eigen::Vector3f KannalaBrandt8::unprojectEig(const cv::Point2f &p2D)"),
    ("src/CameraModels/KannalaBrandt8.cpp", 116, 143, "cv::Point3f KannalaBrandt8::unproject(const cv::Point2f &p2D)"),
    ("src/CameraModels/KannalaBrandt8.cpp", 306, 375, "float KannalaBrandt8::TriangulateMatches("),
    ("src/CameraModels/KannalaBrandt8.cpp", 394, 406, "void KannalaBrandt8::Triangulate("),
]


def main(reference, out):
    parts = ["// GENERATED at build time by oracle/ref_extract_fns.py from %s -- do not commit\n" % reference,
             "namespace ORB_SLAM3 {\n"]
    for rel, a, b, sig in RANGES:
        lines = open(os.path.join(reference, rel), encoding="utf-8", errors="replace").read().split("\n")
        if sig not in lines[a - 1]:
            sys.exit("ref_extract_fns: %s:%d does not start with `%s` (another reference revision?)" % (rel, a, sig))
        if lines[b - 1].strip() != "}":
            sys.exit("ref_extract_fns: %s:%d is not the closing brace of the function" % (rel, b))
        parts.append("// ---- %s:%d-%d\n#line %d \"%s\"\n" % (rel, a, b, a, os.path.join(reference, rel)))
        parts.append("\n".join(lines[a - 1:b]) + "\n")
    parts.append("}  // namespace ORB_SLAM3\n")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write("".join(parts))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
