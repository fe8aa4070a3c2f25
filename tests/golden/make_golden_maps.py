"""Ground-truth maps of the reference's own retrieval fixture, Flickr30k/ann_file/flickr30k_{test,val}.json.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_maps.py
The loop below is the map construction of data/flickr30k_dataset.py:110-118 (flickr30k_retrieval_eval.__init__), which
cannot be instantiated offline (download_url, image files): txt ids are assigned in annotation order, image by image.
Only the integer maps are stored (no captions, no image names).
"""
import json
import os
import sys

REF = "/root/reference/Flickr30k/ann_file"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flickr30k_maps.json")


def build_maps(annotation):
    txt2img, img2txt = {}, {}
    txt_id = 0
    for img_id, ann in enumerate(annotation):                 # flickr30k_dataset.py:111-118
        img2txt[img_id] = []
        for _caption in ann["caption"]:
            img2txt[img_id].append(txt_id)
            txt2img[txt_id] = img_id
            txt_id += 1
    return txt2img, img2txt


def main():
    out = {}
    for split in ("test", "val"):
        with open(os.path.join(REF, f"flickr30k_{split}.json")) as f:
            ann = json.load(f)
        txt2img, img2txt = build_maps(ann)
        out[split] = {"n_img": len(ann), "n_txt": len(txt2img),
                      "txt2img": [txt2img[t] for t in range(len(txt2img))],
                      "img2txt": [img2txt[i] for i in range(len(ann))]}
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(OUT, {k: (v["n_img"], v["n_txt"]) for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
