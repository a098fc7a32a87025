// sq_stage.h -- host-side staging of the reference's call site, in native code behind the C ABI.
//
// What run_multi_view.py:31-58 (reference) derives for every track before it optimises -- load_pred_object
// (src/utils/tracking_gt_utils.py:145-211) per frame, averaging_T_wos (:59-66), the 20 px border rule of bbox_to_lines
// (src/super_quadric/quadric_helper.py:87-107) -- for ALL tracks of a call in one pass over their rows, written
// straight into the packed arrays the optimiser entry takes.  Plain C++ (no CUDA): index bookkeeping on a few
// thousand rows; the numpy form of the same (odam_b200.run_multi_view.stage_tracks_numpy) is kept as its mirror.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/odam_sq.h"

extern "C" int odam_sq_stage_tracks_host(const double *const *tracks, const int64_t *rows_per, int n, int row_stride,
                                         const int64_t *frame_ids, int n_frames, int img_h, int img_w,
                                         int32_t *cls, double *t_wo, double *yaw, double *dims, int32_t *view_off,
                                         int64_t *frame_idx, float *box, uint8_t *mask, int64_t *n_present)
{
    if (n < 0 || n_frames < 0 || row_stride < 13 || !view_off) return ODAM_SQ_ERR_ARG;
    if (n > 0 && (!tracks || !rows_per || !cls || !t_wo || !yaw || !dims || !n_present)) return ODAM_SQ_ERR_ARG;
    if (n_frames > 0 && !frame_ids) return ODAM_SQ_ERR_ARG;
    // frame id -> index into frame_ids (the first index among equal ids, as a stable argsort + leftmost search gives)
    std::vector<std::pair<int64_t, int>> fs((size_t)n_frames);
    for (int f = 0; f < n_frames; f++) fs[f] = {frame_ids[f], f};
    std::stable_sort(fs.begin(), fs.end(), [](const std::pair<int64_t, int> &a, const std::pair<int64_t, int> &b) { return a.first < b.first; });
    std::vector<int64_t> first_row((size_t)n_frames);
    std::vector<double> col;
    const double lim = 20.0;   // EDGE_THRESHOLD, tracking_gt_utils.py:199
    const double hi[4] = {(double)img_w - lim, (double)img_w - lim, (double)img_h - lim, (double)img_h - lim};
    int64_t sv = 0;
    view_off[0] = 0;
    for (int i = 0; i < n; i++) {
        const int64_t R = rows_per[i];
        const double *T = tracks[i];
        if (R < 0 || (R > 0 && !T)) return ODAM_SQ_ERR_ARG;
        // class: int(np.median(column 1)), tracking_gt_utils.py:153
        cls[i] = 0;
        t_wo[3 * i] = t_wo[3 * i + 1] = t_wo[3 * i + 2] = 0.0;
        if (R > 0) {
            col.resize((size_t)R);
            for (int64_t r = 0; r < R; r++) col[r] = T[r * row_stride + 1];
            std::sort(col.begin(), col.end());
            cls[i] = (int32_t)(int64_t)((col[(R - 1) / 2] + col[R / 2]) / 2);
            // centre: mean over ALL rows (:155), accumulated in row order
            double s0 = 0, s1 = 0, s2 = 0;
            for (int64_t r = 0; r < R; r++) { s0 += T[r * row_stride + 9]; s1 += T[r * row_stride + 10]; s2 += T[r * row_stride + 11]; }
            t_wo[3 * i] = s0 / (double)R; t_wo[3 * i + 1] = s1 / (double)R; t_wo[3 * i + 2] = s2 / (double)R;
        }
        // the first row of the track for every frame of frame_ids it contains (np.where(...)[0][0], :181)
        std::fill(first_row.begin(), first_row.end(), (int64_t)-1);
        for (int64_t r = 0; r < R; r++) {
            const int64_t fr = (int64_t)(int32_t)T[r * row_stride];
            auto it = std::lower_bound(fs.begin(), fs.end(), fr, [](const std::pair<int64_t, int> &a, int64_t v) { return a.first < v; });
            if (it == fs.end() || it->first != fr) continue;
            if (first_row[it->second] < 0) first_row[it->second] = r;
        }
        // frames in the order of frame_ids: chordal mean of the yaws, mean dims, boxes with the border rule
        double ss = 0, cs = 0, d0 = 0, d1 = 0, d2 = 0;
        int64_t present = 0;
        for (int f = 0; f < n_frames; f++) {
            const int64_t r = first_row[f];
            if (r < 0) continue;
            const double *row = T + r * row_stride;
            present++;
            ss += sin(row[12]); cs += cos(row[12]);
            d0 += row[6]; d1 += row[7]; d2 += row[8];
            const double b[4] = {row[2], row[4], row[3], row[5]};   // x_min, x_max, y_min, y_max
            bool any = false, m[4];
            for (int k = 0; k < 4; k++) { m[k] = b[k] > lim && b[k] < hi[k]; any |= m[k]; }   // quadric_helper.py:87-107
            if (!any) continue;                                                                // run_multi_view.py:52-55
            if (frame_idx) frame_idx[sv] = f;
            if (box) for (int k = 0; k < 4; k++) box[4 * sv + k] = m[k] ? (float)b[k] : 0.f;
            if (mask) for (int k = 0; k < 4; k++) mask[4 * sv + k] = m[k] ? 1 : 0;
            sv++;
        }
        n_present[i] = present;
        yaw[i] = atan2(ss, cs);
        dims[3 * i] = dims[3 * i + 1] = dims[3 * i + 2] = 0.0;
        if (present) { dims[3 * i] = d0 / (double)present; dims[3 * i + 1] = d1 / (double)present; dims[3 * i + 2] = d2 / (double)present; }
        if (sv > INT32_MAX) return ODAM_SQ_ERR_ARG;
        view_off[i + 1] = (int32_t)sv;
    }
    return ODAM_SQ_OK;
}
