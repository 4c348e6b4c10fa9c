"""String half of the reference's SVG post-processing (starvector/data/util.py:123-136, svg_validator_base.py:380-408)."""
import pytest

from starvector_b200.postprocess import VOID_SVG, clean_svg, is_valid_svg, post_process_svg, process_and_rasterize_svg, process_svg

GOOD = '<svg xmlns="http://www.w3.org/2000/svg" viewBox="0 0 24 24"><path d="M12 2L2 7l10 5 10-5z"/><g><circle cx="1" cy="2" r="3"/></g></svg>'


def test_valid_svg_passes_through_untouched():
    r = post_process_svg(GOOD)
    assert r == {"svg": GOOD, "svg_raw": GOOD, "post_processed": False, "no_compile": False}
    assert process_svg(GOOD) == GOOD


def test_truncated_generation_is_closed():
    cut = '<svg viewBox="0 0 24 24"><g fill="red"><path d="M1 1h2"/><rect x="1" y="2" width="3" hei'      # max_length hit mid-attribute
    r = post_process_svg(cut)
    assert r["post_processed"] and not r["no_compile"] and is_valid_svg(r["svg"])
    assert r["svg"] == '<svg viewBox="0 0 24 24"><g fill="red"><path d="M1 1h2"/></g></svg>' and r["svg_raw"] == cut


def test_trailing_tokens_after_the_root_and_xml_header_are_dropped():
    text = '<?xml version="1.0"?>\n' + GOOD + "<svg><path d='"
    assert clean_svg(text) == GOOD


def test_unrepairable_text_becomes_the_placeholder():
    r = post_process_svg("no markup at all")
    assert r == {"svg": VOID_SVG, "svg_raw": "no markup at all", "post_processed": True, "no_compile": True}


def test_rasterisation_fails_loudly_without_cairosvg():
    try:
        import cairosvg  # noqa: F401
    except Exception:
        with pytest.raises(RuntimeError, match="cairosvg"):
            process_and_rasterize_svg(GOOD)
