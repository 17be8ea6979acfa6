"""TEST INFRASTRUCTURE: numpy restatement of the 2-D overlay image person_msgs/scripts/pose2D_plot_node.py publishes
(draw_humans :18-66, callback_pose :82-91). OpenCV is absent from this image, so cv2.circle / cv2.line / cv2.rectangle
are replaced by the coverage rules csrc/kernels_overlay.cu documents (parity with OpenCV's rasteriser is unpinned; what
IS pinned to the reference are the drawing decisions: thresholds, rounding, colours, sizes, order)."""
import numpy as np

# pose2D_plot_node.py:10-16
COCO_COLORS = [(255, 0, 0), (255, 85, 0), (255, 170, 0), (255, 255, 0), (170, 255, 0), (85, 255, 0), (0, 255, 0),
               (0, 255, 85), (0, 255, 170), (0, 255, 255), (0, 170, 255), (0, 85, 255), (0, 0, 255), (50, 0, 255),
               (100, 0, 255), (170, 0, 255), (255, 0, 255), (255, 150, 0), (85, 170, 0), (42, 128, 85), (0, 85, 170),
               (255, 0, 170), (255, 0, 85), (242, 165, 65)]
COCO_PAIRS = [(0, 1), (0, 2), (1, 3), (2, 4), (3, 5), (4, 6), (5, 7), (6, 8), (7, 9), (8, 10), (5, 11), (6, 12), (11, 13),
              (12, 14), (13, 15), (14, 16)]
CONF_THRESHOLD_DRAW = 0.25   # :19


def _int(v):
    """Python int(x + 0.5) on a float32 value, as the node computes pixel centres (:42, :58-61)."""
    return int(np.float32(v) + np.float32(0.5))


def draw_humans(width, height, persons):
    """persons: structured array of Person2D records. Returns the rgb8 image [height][width][3]."""
    img = np.full((height, width, 3), 255, np.uint8)                       # :85
    yy, xx = np.mgrid[0:height, 0:width].astype(np.int64)
    scale = max(1, int(width / 360))
    r, t_line, h_box = scale * 5, scale * 4, scale
    for ps in persons:
        centers = {}
        for i in range(17):                                                # :33-47
            kp = ps["keypoints"][i]
            if kp["score"] < CONF_THRESHOLD_DRAW:
                continue
            c = (_int(kp["x"]), _int(kp["y"]))
            centers[i] = c
            img[(xx - c[0]) ** 2 + (yy - c[1]) ** 2 <= r * r] = COCO_COLORS[i]
        for a, b in COCO_PAIRS:                                            # :50-54
            if a not in centers or b not in centers:
                continue
            (ax, ay), (bx, by) = centers[a], centers[b]
            dx, dy = bx - ax, by - ay
            px, py = xx - ax, yy - ay
            L2 = dx * dx + dy * dy
            u = px * dx + py * dy
            d_a = 4 * (px * px + py * py) <= t_line * t_line
            d_b = 4 * ((xx - bx) ** 2 + (yy - by) ** 2) <= t_line * t_line
            if L2 == 0:
                hit = d_a
            else:
                mid = 4 * ((px * px + py * py) * L2 - u * u) <= t_line * t_line * L2
                hit = np.where(u <= 0, d_a, np.where(u >= L2, d_b, mid))
            img[hit] = COCO_COLORS[b]
        x1, y1, x2, y2 = (_int(ps["bbox"][0]) - 6, _int(ps["bbox"][1]) - 6, _int(ps["bbox"][2]) + 6,
                          _int(ps["bbox"][3]) + 6)                         # :57-61
        outer = (xx >= x1 - h_box) & (xx <= x2 + h_box) & (yy >= y1 - h_box) & (yy <= y2 + h_box)
        inner = (xx >= x1 + h_box) & (xx <= x2 - h_box) & (yy >= y1 + h_box) & (yy <= y2 - h_box)
        img[outer & ~inner] = COCO_COLORS[0]                               # id = 0 (:84), colors[id % len]
    return img
