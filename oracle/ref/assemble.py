"""oracle/ref/assemble.py — TEST INFRASTRUCTURE ONLY.

Assembles the reference's fragment shaders the way its JavaScript does at run time (string replaces, #include
resolution, `unrollLoops`, #define prefix), reading the shader text from the reference checkout
(RFX_REFERENCE_DIR, default /root/reference).  The JS it mirrors is cited per function.  The result is handed to
oracle/ref/transpile.py; nothing read here is written into the repository.

Third-party text the reference pulls in at run time and that is NOT in the checkout — three.js shader chunks (`three`
0.151.3, package-lock.json:6303) and the postprocessing `Effect` wrapper (`postprocessing` ^6.30) — is restated below
from the published sources, limited to what the hot-path shaders reference.
"""
from __future__ import annotations

import os
import re

REF = os.environ.get("RFX_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "ssgi", "shader", "ssgi.frag"))


def read(rel: str) -> str:
    with open(os.path.join(REF, "src", rel), encoding="utf-8") as f:
        return f.read().replace("﻿", "")


# ------------------------------------------------------------------------------------------------ three.js chunks
# three r151 src/renderers/shaders/ShaderChunk/common.glsl.js (subset: constants, helpers the shaders can see)
THREE_COMMON = """
#define PI 3.141592653589793
#define PI2 6.283185307179586
#define PI_HALF 1.5707963267948966
#define RECIPROCAL_PI 0.3183098861837907
#define RECIPROCAL_PI2 0.15915494309189535
#define EPSILON 1e-6
#ifndef saturate
#define saturate( a ) clamp( a, 0.0, 1.0 )
#endif
#define whiteComplement( a ) ( 1.0 - saturate( a ) )
float pow2( const in float x ) { return x*x; }
vec3 pow2( const in vec3 x ) { return x*x; }
float pow3( const in float x ) { return x*x*x; }
float pow4( const in float x ) { float x2 = x*x; return x2*x2; }
float max3( const in vec3 v ) { return max( max( v.x, v.y ), v.z ); }
float average( const in vec3 v ) { return dot( v, vec3( 0.3333333 ) ); }
float luminance( const in vec3 rgb ) {
	const vec3 weights = vec3( 0.2126729, 0.7151522, 0.0721750 );
	return dot( weights, rgb );
}
highp float rand( const in vec2 uv ) {
	const highp float a = 12.9898, b = 78.233, c = 43758.5453;
	highp float dt = dot( uv.xy, vec2( a,b ) ), sn = mod( dt, PI );
	return fract( sin( sn ) * c );
}
"""
# three r151 ShaderChunk/encodings_pars_fragment.glsl.js: what `linearToOutputTexel` resolves to for an sRGB output (WebGLProgram
# getTexelEncodingFunction) and for a linear one
THREE_LINEAR_TO_OUTPUT = {
    True: """
vec4 LinearTosRGB( in vec4 value ) {
	return vec4( mix( pow( value.rgb, vec3( 0.41666 ) ) * 1.055 - vec3( 0.055 ), value.rgb * 12.92, vec3( lessThanEqual( value.rgb, vec3( 0.0031308 ) ) ) ), value.a );
}
vec4 linearToOutputTexel( vec4 value ) { return LinearTosRGB( value ); }
""",
    False: """
vec4 linearToOutputTexel( vec4 value ) { return value; }
""",
}
# three r151 ShaderChunk/packing.glsl.js (subset)
THREE_PACKING = """
vec3 packNormalToRGB( const in vec3 normal ) {
	return normalize( normal ) * 0.5 + 0.5;
}
vec3 unpackRGBToNormal( const in vec3 rgb ) {
	return 2.0 * rgb.xyz - 1.0;
}
float viewZToOrthographicDepth( const in float viewZ, const in float near, const in float far ) {
	return ( viewZ + near ) / ( near - far );
}
float orthographicDepthToViewZ( const in float depth, const in float near, const in float far ) {
	return depth * ( near - far ) - near;
}
float viewZToPerspectiveDepth( const in float viewZ, const in float near, const in float far ) {
	return ( ( near + viewZ ) * far ) / ( ( far - near ) * viewZ );
}
float perspectiveDepthToViewZ( const in float depth, const in float near, const in float far ) {
	return ( near * far ) / ( ( far - near ) * depth - far );
}
"""
THREE_FOG_PARS_FRAGMENT = """
#ifdef USE_FOG
	uniform vec3 fogColor;
	varying float vFogDepth;
	#ifdef FOG_EXP2
		uniform float fogDensity;
	#else
		uniform float fogNear;
		uniform float fogFar;
	#endif
#endif
"""
THREE_FOG_FRAGMENT = """
#ifdef USE_FOG
	#ifdef FOG_EXP2
		float fogFactor = 1.0 - exp( - fogDensity * fogDensity * vFogDepth * vFogDepth );
	#else
		float fogFactor = smoothstep( fogNear, fogFar, vFogDepth );
	#endif
	gl_FragColor.rgb = mix( gl_FragColor.rgb, fogColor, fogFactor );
#endif
"""
CHUNKS = {"common": THREE_COMMON, "packing": THREE_PACKING, "fog_pars_fragment": THREE_FOG_PARS_FRAGMENT, "fog_fragment": THREE_FOG_FRAGMENT}

# what WebGLProgram prepends to a ShaderMaterial fragment shader (three r151 WebGLProgram.js prefixFragment), minus
# tone mapping / colour-space helpers the passes do not call (toneMapped: false)
THREE_FRAGMENT_PREFIX = """
uniform mat4 viewMatrix;
uniform vec3 cameraPosition;
uniform bool isOrthographic;
"""


def resolve_includes(src: str) -> str:
    """three r151 WebGLProgram.js resolveIncludes: every `#include <name>` line, recursively."""

    def rep(m):
        name = m.group(1)
        if name not in CHUNKS:
            raise KeyError(f"#include <{name}>: not a three.js chunk (three's resolver throws here too)")
        return resolve_includes(CHUNKS[name])

    return re.sub(r"^[ \t]*#include +<([\w\d./]+)>", rep, src, flags=re.M)


def unroll_loops(src: str) -> str:
    """src/ssgi/utils/Utils.js:74-89"""
    pat = re.compile(r"#pragma unroll_loop_start\s+for\s*\(\s*int\s+i\s*=\s*(\d+)\s*;\s*i\s*<\s*(\d+)\s*;\s*i\s*\+\+\s*\)\s*{([\s\S]+?)}\s+#pragma unroll_loop_end")

    def rep(m):
        out = ""
        for i in range(int(m.group(1)), int(m.group(2))):
            out += re.sub(r"\[\s*i\s*\]", f"[ {i} ]", m.group(3)).replace("UNROLLED_LOOP_INDEX", str(i))
        return out

    return pat.sub(rep, src)


def setup_blue_noise(fragment: str) -> str:
    """src/utils/BlueNoiseUtils.js:17-37 (the shader-text part)"""
    return fragment.replace("uniform vec2 resolution;", "uniform vec2 resolution;\n" + read("utils/shader/blue_noise.glsl"), 1)


def js_to_precision5(v: float) -> str:
    """Number.prototype.toPrecision(5) for the magnitudes the options take"""
    return "%#.5g" % v if v != 0 else "0.0000"


def finish(fragment: str, defines: dict) -> str:
    """WebGLProgram: #define lines (generateDefines), prefix, resolved includes."""
    head = "".join(f"#define {k} {v}\n" for k, v in defines.items() if v is not False)
    return head + THREE_FRAGMENT_PREFIX + resolve_includes(fragment)


# ------------------------------------------------------------------------------------------------ the passes
def ssgi(*, steps=20, refine_steps=5, mode=0, importance_sampling=True, missed_rays=False, use_direct_light=True, use_envmap=True,
         perspective=True) -> str:
    """src/ssgi/material/SSGIMaterial.js:44-58 + useBlueNoise; defines from SSGIPass.js:40-42, SSGIEffect.js:129-135,203-221,334-346"""
    frag = read("ssgi/shader/ssgi.frag")
    frag = frag.replace("#include <ssgi_utils>", read("ssgi/shader/ssgi_utils.frag"), 1)
    frag = frag.replace("#include <gbuffer_packing>", read("gbuffer/shader/gbuffer_packing.glsl"), 1)
    frag = setup_blue_noise(frag)
    d = {"steps": int(steps), "refineSteps": int(refine_steps), "CUBEUV_TEXEL_WIDTH": 0, "CUBEUV_TEXEL_HEIGHT": 0, "CUBEUV_MAX_MIP": 0,
         "vWorldPosition": "worldPos"}
    if perspective:
        d["PERSPECTIVE_CAMERA"] = ""
    d["mode"] = int(mode)
    if use_envmap:
        d["USE_ENVMAP"] = ""
        if importance_sampling:
            d["importanceSampling"] = ""
    if missed_rays:
        d["missedRays"] = ""
    if use_direct_light:
        d["useDirectLight"] = ""
    return finish(frag, d)


def temporal_reproject(*, texture_count=2, input_type=0, dilation=False, neighborhood_clamp=(False, True), reproject_specular=(False, True),
                       log_transform=False, neighborhood_clamp_radius=1, depth_distance=2.0, world_distance=4.0, confidence_power=0.75,
                       perspective=True) -> str:
    """src/temporal-reproject/material/TemporalReprojectMaterial.js:12-43 + TemporalReprojectPass.js:76-117"""
    frag = read("temporal-reproject/shader/temporal_reproject.frag")
    frag = frag.replace("#include <reproject>", read("temporal-reproject/shader/reproject.frag"), 1)
    frag = frag.replace("#include <gbuffer_packing>", read("gbuffer/shader/gbuffer_packing.glsl"), 1)
    definitions = ""
    for i in range(texture_count):
        definitions += f"\n\t\t\t\tuniform sampler2D accumulatedTexture{i};\n\n\t\t\t\tlayout(location = {i}) out vec4 gOutput{i};\n\t\t\t"
    frag = definitions + frag.replace("textureCount", str(texture_count))
    frag = unroll_loops(frag)
    frag = re.sub(r"accumulatedTexture\[\s*([0-9]+)\s*]", r"accumulatedTexture\1", frag)
    frag = re.sub(r"gOutput\[\s*([0-9]+)\s*]", r"gOutput\1", frag)

    def arr(v):
        # TemporalReprojectPass.js:109-115: `typeof value !== "array"` is always true in JS, so even an array option is wrapped:
        # Array(textureCount).fill(value).join(", ") — an array value stringifies as "a,b" inside each slot
        one = ",".join("true" if x else "false" for x in v) if isinstance(v, (tuple, list)) else ("true" if v else "false")
        return "bool[](" + ", ".join([one] * texture_count) + ")"

    d = {"textureCount": texture_count}
    if dilation:
        d["dilation"] = ""
    if any(neighborhood_clamp) if isinstance(neighborhood_clamp, (tuple, list)) else neighborhood_clamp:
        d["neighborhoodClamp"] = ""  # overwritten below by the bool[] form (TemporalReprojectPass.js:109-115), as in the JS
    if log_transform:
        d["logTransform"] = ""
    if perspective:
        d["PERSPECTIVE_CAMERA"] = ""
    d["neighborhoodClampRadius"] = int(neighborhood_clamp_radius)
    d["depthDistance"] = js_to_precision5(depth_distance)
    d["worldDistance"] = js_to_precision5(world_distance)
    d["inputType"] = int(input_type)
    d["reprojectSpecular"] = arr(reproject_specular)
    d["neighborhoodClamp"] = arr(neighborhood_clamp)
    d["confidencePower"] = js_to_precision5(confidence_power)
    return finish(frag, d)


def poisson_denoise(*, input_type="diffuseSpecular", gbuffer=True) -> str:
    """src/denoise/pass/PoissonDenoisePass.js:16,37-72,108-117"""
    frag = read("denoise/shader/poisson_denoise.frag").replace("#include <gbuffer_packing>", read("gbuffer/shader/gbuffer_packing.glsl"), 1)
    is_spec = {"diffuseSpecular": [False, True], "diffuse": [False, False], "specular": [True, True]}[input_type]
    texture_count = 2 if input_type == "diffuseSpecular" else 1
    frag = unroll_loops(frag.replace("textureCount", str(texture_count)))
    frag = setup_blue_noise(frag)
    d = {"isTextureSpecular": "bool[2](" + ",".join("true" if x else "false" for x in is_spec) + ")"}
    if gbuffer:
        d["GBUFFER_TEXTURE"] = ""
    else:
        # REPAIR (the only edit of reference text in this file).  The branch without GBUFFER_TEXTURE — the one AOEffect's denoiser
        # uses — does not compile as GLSL at this commit: poisson_denoise.frag:58-60 swizzles a texel to .xyz into a vec3 and then
        # reads `.a` of it (HBAOEffect does not run in the reference).  The minimal repair keeps the texel a vec4;
        # results of this branch are "reference + this repair", and the oracle restates the same repaired form.
        broken = "vec3 neighborDepthVelocityTexel = textureLod(normalTexture, neighborUv, 0.).xyz;"
        assert broken in frag
        frag = frag.replace(broken, "vec4 neighborDepthVelocityTexel = textureLod(normalTexture, neighborUv, 0.);", 1)
    return finish(frag, d)


def gi_compose(*, input_type=0, perspective=True) -> str:
    """src/denoise/pass/DenoiserComposePass.js:36-86 (the template literal) + :100-110"""
    js = read("denoise/pass/DenoiserComposePass.js")
    m = re.search(r"fragmentShader:\s*/\*\s*glsl\s*\*/\s*`(.*?)`", js, flags=re.S)
    frag = m.group(1)
    frag = frag.replace("${gbuffer_packing}", read("gbuffer/shader/gbuffer_packing.glsl"))
    frag = frag.replace("${ssgi_poisson_compose_functions}", read("denoise/shader/denoiser_compose_functions.glsl"))
    d = {"inputType": int(input_type)}
    if perspective:
        d["PERSPECTIVE_CAMERA"] = ""
    return finish(frag, d)


def hbao(*, spp=8, animated_noise=False, use_normal_texture=False) -> str:
    """src/hbao/HBAOPass.js:6 + src/ao/AOPass.js:30-66 (useBlueNoise); `#include <sampleBlueNoise>` (hbao_utils.glsl:1) is not a
    three.js chunk — three's resolver throws on it; it is dropped, as the oracle does"""
    frag = read("hbao/shader/hbao.frag").replace("#include <hbao_utils>", read("hbao/shader/hbao_utils.glsl"), 1)
    frag = frag.replace("#include <sampleBlueNoise>", "", 1)
    frag = setup_blue_noise(frag)
    d = {"spp": int(spp)}  # AOEffect.js:76-80
    if animated_noise:
        d["animatedNoise"] = ""
    if use_normal_texture:
        d["useNormalTexture"] = ""
    return finish(frag, d)


# postprocessing's EffectMaterial: the effect's mainImage() is called with the input buffer's texel and vUv; the result is
# written with the effect's blend function (NORMAL: dst = src) — postprocessing 6.x src/materials/glsl/effect.frag
EFFECT_HEAD = ("varying vec2 vUv;\nuniform sampler2D inputBuffer;\nuniform vec2 resolution;\nuniform vec2 texelSize;\nuniform float cameraNear;\nuniform float cameraFar;\n"
               + THREE_COMMON + THREE_PACKING)
EFFECT_MAIN = """
void main() {
  vec4 color0 = texture2D(inputBuffer, vUv);
  vec4 color1 = vec4(0.0);
  mainImage(color0, vUv, color1);
  gl_FragColor = color1;
}
"""


def _effect(frag: str, defines: dict) -> str:
    return finish(EFFECT_HEAD + "#define texture2D(s, uv) textureLod(s, uv, 0.0)\n" + frag + EFFECT_MAIN, defines)


def motion_blur(*, samples=16) -> str:
    """src/motion-blur/MotionBlurEffect.js:22-46"""
    frag = setup_blue_noise(read("motion-blur/shader/motion_blur.frag"))
    s = "%d" % round(samples)
    return _effect(frag, {"samples": s, "samplesFloat": s + ".0"})


def ao_compose() -> str:
    """src/ao/AOEffect.js:27-35"""
    return _effect(read("ao/shader/ao_compose.frag"), {})


def traa_compose() -> str:
    """src/traa/TRAAEffect.js:11-15"""
    return _effect(read("traa/shader/traa_compose.frag"), {})


def ssgi_compose(*, fog=False, fog_exp2=False, perspective=True) -> str:
    """src/ssgi/SSGIEffect.js:34-64"""
    frag = read("ssgi/shader/ssgi_compose.frag")
    frag = frag.replace("#include <fog_pars_fragment>", THREE_FOG_PARS_FRAGMENT.replace("varying", "", 1), 1)
    frag = frag.replace("#include <fog_fragment>", re.sub(r".*gl_FragColor.*", "", THREE_FOG_FRAGMENT), 1)
    d = {"PERSPECTIVE_CAMERA": "1" if perspective else "0"}
    if fog:
        d["USE_FOG"] = ""
    if fog_exp2:
        d["FOG_EXP2"] = ""
    return _effect(frag, d)


def _js_template(rel: str) -> str:
    """the `const fragmentShader = /* glsl */ `...`` literal of an Effect's JS file"""
    m = re.search(r"const fragmentShader\s*=\s*/\*\s*glsl\s*\*/\s*`(.*?)`", read(rel), flags=re.S)
    return m.group(1)


def sharpness() -> str:
    """src/sharpness/SharpnessEffect.js:4-30"""
    return _effect(_js_template("sharpness/SharpnessEffect.js"), {})


def lens_distortion() -> str:
    """src/lens-distortion/LensDistortionEffect.js:5-46"""
    return _effect(_js_template("lens-distortion/LensDistortionEffect.js"), {})


def gradual_background(*, perspective=True) -> str:
    """src/gradual-background/GradualBackgroundEffect.js:3-47,59"""
    return _effect(_js_template("gradual-background/GradualBackgroundEffect.js"), {"PERSPECTIVE_CAMERA": "1" if perspective else "0"})


def sparkle(*, perspective=True) -> str:
    """src/sparkle/SparkleEffect.js:4-100.  PERSPECTIVE_CAMERA is never defined by SparkleEffect (its `#if PERSPECTIVE_CAMERA == 1` therefore
    takes the orthographic branch in the reference); `perspective` selects what a host that defines it gets."""
    frag = _js_template("sparkle/SparkleEffect.js").replace("${gbuffer_packing}", read("gbuffer/shader/gbuffer_packing.glsl"))
    return _effect(frag, {"PERSPECTIVE_CAMERA": "1"} if perspective else {})


def taa(*, srgb_output=True) -> str:
    """src/taa/TAAPass.js:37-50 + shader/taa.frag; rendered to the screen, so linearToOutputTexel is the renderer's output transfer"""
    return finish(THREE_LINEAR_TO_OUTPUT[bool(srgb_output)] + read("taa/shader/taa.frag"), {})
