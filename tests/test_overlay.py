"""the 2-D overlay renderer (person_msgs/scripts/pose2D_plot_node.py) on the GPU against the numpy
restatement oracle/overlay_ref.py, pixel for pixel, plus properties that follow from the node's drawing decisions."""
import numpy as np
import pytest

from oracle import overlay_ref
from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype
from tests import helpers


def test_overlay_oracle_follows_the_nodes_drawing_decisions():
    ps = np.zeros(1, person2d_dtype)
    ps["keypoints"]["x"][0, :] = np.linspace(100, 500, 17)
    ps["keypoints"]["y"][0, :] = 200
    ps["keypoints"]["score"][0, :] = 0.9
    ps["keypoints"]["score"][0, 3] = 0.2          # below _CONF_THRESHOLD_DRAW: no circle, no limb through it
    ps["bbox"][0] = (100, 180, 500, 220)
    img = overlay_ref.draw_humans(640, 480, ps)
    assert tuple(img[0, 0]) == (255, 255, 255)
    for k in range(17):
        c = (int(ps["keypoints"]["x"][0, k] + 0.5), 200)
        if k == 3:
            assert tuple(img[c[1] + 4, c[0]]) == (255, 255, 255)
        else:
            assert tuple(img[c[1] + 4, c[0]]) in {overlay_ref.COCO_COLORS[k]} | {overlay_ref.COCO_COLORS[b] for a, b in overlay_ref.COCO_PAIRS}
    # bounding box: 6 px outside the bbox, colour 0, 2 px thick at 640 wide
    assert tuple(img[180 - 6, 300]) == overlay_ref.COCO_COLORS[0] and tuple(img[180 - 6 - 2, 300]) == (255, 255, 255)
    # sizes scale with the width like max(1, int(w / 360))
    big = overlay_ref.draw_humans(1280, 480, ps)
    assert (big != 255).any(axis=-1).sum() > (img != 255).any(axis=-1).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(640, 480), (1280, 720), (333, 251)])
def test_overlay_matches_the_oracle_pixel_for_pixel(size):
    from smartedgesensor3dhumanpose_b200 import api
    w, h = size
    fr = helpers.make_workload("cfg2_hall16x6", 3)
    pipe = api.GeometryPipeline(fr["cameras"])
    persons = fr["persons"].reshape(-1, fr["persons"].shape[2]).copy()
    n_persons = fr["n_persons"].reshape(-1).copy()
    if (w, h) != (1280, 720):   # detections were generated for 1280 x 720 images: squeeze them into the smaller canvas
        persons["keypoints"]["x"] *= w / 1280.0
        persons["keypoints"]["y"] *= h / 720.0
        persons["bbox"][..., 0::2] *= w / 1280.0
        persons["bbox"][..., 1::2] *= h / 720.0
    persons["keypoints"]["score"][0, 0, :5] = 0.1      # joints below the drawing threshold
    persons["keypoints"]["x"][1, 0, 0] = -40.0          # partly outside the canvas
    persons["bbox"][2, 0] = (-20, -20, w + 30, h + 30)  # box larger than the image
    sel = [i for i in range(len(n_persons)) if n_persons[i] > 0][:6] + [int(np.argmin(n_persons))]
    got = pipe.overlay_batch(persons[sel], n_persons[sel], w, h)
    assert got.shape == (len(sel), h, w, 3)
    for j, i in enumerate(sel):
        want = overlay_ref.draw_humans(w, h, persons[i, :n_persons[i]])
        assert np.array_equal(got[j], want), f"image {i}: {(got[j] != want).any(-1).sum()} pixels differ"
    assert (got[:-1] != 255).any()
    pipe.close()
