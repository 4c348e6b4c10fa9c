"""SVG post-processing after generation: the string half of the reference's
`process_and_rasterize_svg` (starvector/data/util.py:123-136) and of the validators' `post_process_svg`
(starvector/validation/svg_validator_base.py:380-408):

    try the raw text -> else `clean_svg(text)` -> else the empty placeholder `<svg></svg>` (util.py:22,120-121).

The reference decides "valid" with `svgpathtools.svgstr2paths` and repairs with BeautifulSoup('xml').prettify() +
`cairosvg.svg2svg` (util.py:88-117).  None of the three packages exists in this image, so:
  * when they ARE importable (a deployment with the reference's requirements installed) they are used exactly as the
    reference uses them;
  * otherwise validity = "parses as XML with an <svg> root" (xml.etree) and the repair is a small, dependency-free one:
    cut everything after the last complete tag, close the elements still open, drop `<?xml` header lines like the
    reference does (util.py:116).  This is a stand-in, not a restatement: the two repairs agree on well-formed input and on
    truncated generations (the common failure: `max_length` hit before `</svg>`), not byte for byte on arbitrary garbage.
Rasterisation (`rasterize_svg`, util.py:138+) needs cairosvg and raises a clear error without it.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from typing import Dict, Tuple

VOID_SVG = "<svg></svg>"                                   # util.py:22


def use_placeholder() -> str:                               # util.py:120-121
    return VOID_SVG


def _have(mod: str) -> bool:
    try:
        __import__(mod)
        return True
    except Exception:                                       # noqa: BLE001
        return False


def is_valid_svg(text: str) -> bool:
    """`svgstr2paths(text)` does not raise (reference) / well-formed XML with an <svg> root (stand-in)."""
    if _have("svgpathtools"):
        from svgpathtools import svgstr2paths

        try:
            svgstr2paths(text)
            return True
        except Exception:                                   # noqa: BLE001 - the reference catches everything too
            return False
    try:
        root = ET.fromstring(text)
    except ET.ParseError:
        return False
    return root.tag.split("}")[-1] == "svg"


_TAG = re.compile(r"<(/?)([A-Za-z_][\w:.-]*)((?:[^<>\"']|\"[^\"]*\"|'[^']*')*?)(/?)>")


def clean_svg(svg_text: str) -> str:
    """util.py:88-117 when bs4 + cairosvg are installed, else the dependency-free repair described in the module docstring."""
    if _have("bs4") and _have("cairosvg"):
        import cairosvg
        from bs4 import BeautifulSoup

        svg_bs4 = BeautifulSoup(svg_text, "xml").prettify()
        svg_cairo = cairosvg.svg2svg(svg_bs4).decode()
        return "\n".join(line for line in svg_cairo.split("\n") if not line.strip().startswith("<?xml"))
    text = "\n".join(line for line in svg_text.split("\n") if not line.strip().startswith("<?xml"))
    start = text.find("<svg")
    if start < 0:
        raise ValueError("no <svg element in the generated text")
    text = text[start:]
    out, stack, pos = [], [], 0
    for m in _TAG.finditer(text):
        closing, name, _attrs, selfclose = m.group(1), m.group(2), m.group(3), m.group(4)
        if closing:
            if name not in stack:
                continue                                    # stray closing tag: drop it
            out.append(text[pos:m.start()])
            while stack and stack[-1] != name:              # close what the generation left open inside
                out.append(f"</{stack.pop()}>")
            stack.pop()
            out.append(m.group(0))
        else:
            out.append(text[pos:m.end()])
            if not selfclose:
                stack.append(name)
        pos = m.end()
        if not stack:
            break                                           # the root <svg> just closed: ignore whatever follows
    while stack:                                            # truncated generation: close the open elements
        out.append(f"</{stack.pop()}>")
    return "".join(out)


def post_process_svg(text: str) -> Dict[str, object]:
    """svg_validator_base.py:380-408: the dict the validators store per sample."""
    if is_valid_svg(text):
        return {"svg": text, "svg_raw": text, "post_processed": False, "no_compile": False}
    try:
        cleaned = clean_svg(text)
        if not is_valid_svg(cleaned):
            raise ValueError("still invalid")
        return {"svg": cleaned, "svg_raw": text, "post_processed": True, "no_compile": False}
    except Exception:                                       # noqa: BLE001
        return {"svg": use_placeholder(), "svg_raw": text, "post_processed": True, "no_compile": True}


def process_svg(svg_string: str) -> str:
    """The SVG string `process_and_rasterize_svg` returns (util.py:123-134), without the raster image."""
    return str(post_process_svg(svg_string)["svg"])


def process_and_rasterize_svg(svg_string: str, resolution: int = 256, dpi: int = 128, scale: int = 2) -> Tuple[str, object]:
    """util.py:123-136.  Needs cairosvg + PIL for the raster half."""
    out_svg = process_svg(svg_string)
    if not _have("cairosvg"):
        raise RuntimeError("rasterisation needs cairosvg, which is not installed in this environment; process_svg() gives the string half")
    import io

    import cairosvg
    from PIL import Image

    png = cairosvg.svg2png(bytestring=out_svg, background_color="white", output_width=resolution, output_height=resolution, dpi=dpi, scale=scale)
    return out_svg, Image.open(io.BytesIO(png))
